// DeepMimic motion-imitation reward, register-resident fast path for the humanoid3d character.
//
// Same arithmetic as imitation.cuh (reference DeepMimicCore scenes/SceneImitate.cpp:7-127 and the KinTree /
// MathUtil / Motion / RBDUtil functions cited there), restructured for the B200 FP32 pipe:
//   * the joint tree of deepmimic/data/characters/humanoid3d.txt (types, parents, parameter offsets, attach
//     points) is a compile-time table, so the tree walk is fully unrolled, every joint
//     state lives in registers under a static name (no local memory) and zero attach components fold away;
//   * world rotations are carried as unit quaternions and expanded to the 3x3 matrix once per joint; the
//     matrix serves the children's attach points, the body attach point and the joint angular velocity;
//   * one persistent 16-warp block per SM; every warp runs its own pipeline over 32-row tiles: lane 0 stages
//     the tile's pose / velocity rows in shared memory with two 1-D TMA bulk copies (cp.async.bulk) on the
//     warp's own mbarrier, so there is no block-wide barrier after the prologue;
//   * the clip tables sit in shared memory once per SM in a 16-byte-slot layout (one float4 per joint) so
//     that a thread fetches a joint of a clip frame with one LDS.128;
//   * acos / sin / cos / exp use the SFU approximations (error << the 1e-3 budget of the reward).
// simstep_load_clip selects this kernel when the loaded character equals the built-in table; any other
// character runs the generic kernel of imitation.cuh.
#pragma once
#include <utility>

#include "imitation.cuh"
#include "ptx.cuh"

namespace simstep {
namespace h3d {

constexpr int kJoints = 15;
constexpr int kDof = 43;
constexpr int kSlots = 11;  // float4 slots per clip row: root pos | root quat | 8 spherical | 4 revolute

// joint types: 0 root, 1 spherical, 2 revolute, 3 fixed
__host__ __device__ constexpr int jtype(int j) {
  constexpr int t[kJoints] = {0, 1, 1, 1, 2, 1, 1, 2, 3, 1, 2, 1, 1, 2, 3};
  return t[j];
}
__host__ __device__ constexpr int jparent(int j) {
  constexpr int t[kJoints] = {-1, 0, 1, 0, 3, 4, 1, 6, 7, 0, 9, 10, 1, 12, 13};
  return t[j];
}
__host__ __device__ constexpr int joff(int j) {
  constexpr int t[kJoints] = {0, 7, 11, 15, 19, 20, 24, 28, 29, 29, 33, 34, 38, 42, 43};
  return t[j];
}
__host__ __device__ constexpr bool jee(int j) { return j == 5 || j == 8 || j == 11 || j == 14; }
// float4 slot of a joint in the permuted clip row; revolute joints share slot 10 (component = jsub)
__host__ __device__ constexpr int jslot(int j) {
  constexpr int t[kJoints] = {1, 2, 3, 4, 10, 5, 6, 10, -1, 7, 10, 8, 9, 10, -1};
  return t[j];
}
__host__ __device__ constexpr int jsub(int j) { return j == 4 ? 0 : j == 7 ? 1 : j == 10 ? 2 : j == 13 ? 3 : 0; }
__host__ __device__ constexpr float jattach(int j, int k) {
  constexpr float t[kJoints][3] = {{0.f, 0.f, 0.f},           {0.f, 0.236151f, 0.f},  {0.f, 0.223894f, 0.f},
                                   {0.f, 0.f, 0.084887f},     {0.f, -0.421546f, 0.f}, {0.f, -0.40987f, 0.f},
                                   {-0.02405f, 0.2435f, 0.18311f}, {0.f, -0.274788f, 0.f}, {0.f, -0.258947f, 0.f},
                                   {0.f, 0.f, -0.084887f},    {0.f, -0.421546f, 0.f}, {0.f, -0.40987f, 0.f},
                                   {-0.02405f, 0.2435f, -0.18311f}, {0.f, -0.274788f, 0.f}, {0.f, -0.258947f, 0.f}};
  return t[j][k];
}
__host__ __device__ constexpr float jbody(int j, int k) {
  constexpr float t[kJoints][3] = {{0.f, 0.07f, 0.f},  {0.f, 0.12f, 0.f},  {0.f, 0.175f, 0.f}, {0.f, -0.21f, 0.f},
                                   {0.f, -0.2f, 0.f},  {0.045f, -0.0225f, 0.f}, {0.f, -0.14f, 0.f}, {0.f, -0.12f, 0.f},
                                   {0.f, 0.f, 0.f},    {0.f, -0.21f, 0.f}, {0.f, -0.2f, 0.f},  {0.045f, -0.0225f, 0.f},
                                   {0.f, -0.14f, 0.f}, {0.f, -0.12f, 0.f}, {0.f, 0.f, 0.f}};
  return t[j][k];
}
__host__ __device__ constexpr bool has_children(int j) {
  for (int i = 0; i < kJoints; ++i)
    if (jparent(i) == j) return true;
  return false;
}

// Does a loaded character equal the built-in table?  (host side, simstep_load_clip)
inline bool matches(const simstep_character* ch) {
  if (ch->n_joints != kJoints || ch->dof != kDof) return false;
  for (int j = 0; j < kJoints; ++j) {
    if (ch->joint_type[j] != jtype(j) || ch->parent[j] != jparent(j) || ch->param_offset[j] != joff(j)) return false;
    if ((ch->is_end_eff[j] != 0) != jee(j)) return false;
    for (int k = 0; k < 3; ++k) {
      if (j > 0 && ch->attach[j][k] != jattach(j, k)) return false;
      if (ch->body_attach[j][k] != jbody(j, k)) return false;
    }
  }
  return true;
}

// Host side: clip row [dof] -> permuted row [kSlots * 4] (unused lanes zero).
inline void permute_row(const float* src, float* dst) {
  for (int i = 0; i < kSlots * 4; ++i) dst[i] = 0.f;
  for (int i = 0; i < 3; ++i) dst[i] = src[i];
  for (int i = 0; i < 4; ++i) dst[4 + i] = src[3 + i];
  for (int j = 1; j < kJoints; ++j) {
    if (jtype(j) == 1)
      for (int i = 0; i < 4; ++i) dst[4 * jslot(j) + i] = src[joff(j) + i];
    else if (jtype(j) == 2)
      dst[4 * jslot(j) + jsub(j)] = src[joff(j)];
  }
}

// ---- small math ----------------------------------------------------------------------------------------

struct Mat3 { float xx, xy, xz, yx, yy, yz, zx, zy, zz; };

// rotation matrix of a UNIT quaternion
__device__ __forceinline__ Mat3 qmat(Quat q) {
  const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
  const float xx = q.x * x2, yy = q.y * y2, zz = q.z * z2;
  const float xy = q.x * y2, xz = q.x * z2, yz = q.y * z2;
  const float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
  Mat3 m;
  m.xx = 1.f - (yy + zz); m.xy = xy - wz;         m.xz = xz + wy;
  m.yx = xy + wz;         m.yy = 1.f - (xx + zz); m.yz = yz - wx;
  m.zx = xz - wy;         m.zy = yz + wx;         m.zz = 1.f - (xx + yy);
  return m;
}
__device__ __forceinline__ Vec3 mul(const Mat3& m, Vec3 v) {
  return v3(fmaf(m.xx, v.x, fmaf(m.xy, v.y, m.xz * v.z)), fmaf(m.yx, v.x, fmaf(m.yy, v.y, m.yz * v.z)),
            fmaf(m.zx, v.x, fmaf(m.zy, v.y, m.zz * v.z)));
}
// m * (compile-time constant vector): zero components cost nothing
template <int J, bool BODY>
__device__ __forceinline__ Vec3 mul_const(const Mat3& m) {
  constexpr float cx = BODY ? jbody(J, 0) : jattach(J, 0);
  constexpr float cy = BODY ? jbody(J, 1) : jattach(J, 1);
  constexpr float cz = BODY ? jbody(J, 2) : jattach(J, 2);
  Vec3 r = v3(0.f, 0.f, 0.f);
  if constexpr (cx != 0.f) { r.x = m.xx * cx; r.y = m.yx * cx; r.z = m.zx * cx; }
  if constexpr (cy != 0.f) {
    if constexpr (cx != 0.f) { r.x = fmaf(m.xy, cy, r.x); r.y = fmaf(m.yy, cy, r.y); r.z = fmaf(m.zy, cy, r.z); }
    else { r.x = m.xy * cy; r.y = m.yy * cy; r.z = m.zy * cy; }
  }
  if constexpr (cz != 0.f) {
    if constexpr (cx != 0.f || cy != 0.f) { r.x = fmaf(m.xz, cz, r.x); r.y = fmaf(m.yz, cz, r.y); r.z = fmaf(m.zz, cz, r.z); }
    else { r.x = m.xz * cz; r.y = m.yz * cz; r.z = m.zz * cz; }
  }
  return r;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// acos on [-1, 1]: sqrt(1-|x|) * P7(|x|) (Abramowitz & Stegun 4.4.46, |error| <= 2e-8), reflected for x < 0
__device__ __forceinline__ float fast_acos(float x) {
  const float a = fabsf(x);
  float p = -0.0012624911f;
  p = fmaf(p, a, 0.0066700901f);
  p = fmaf(p, a, -0.0170881256f);
  p = fmaf(p, a, 0.0308918810f);
  p = fmaf(p, a, -0.0501743046f);
  p = fmaf(p, a, 0.0889789874f);
  p = fmaf(p, a, -0.2145988016f);
  p = fmaf(p, a, 1.5707963050f);
  const float r = p * sqrt_approx(fmaxf(1.f - a, 0.f));
  return x < 0.f ? 3.14159265358979f - r : r;
}

// QuatTheta(QuatDiff(q0, q1))^2, see quat_theta_sq in imitation.cuh
__device__ __forceinline__ float theta_sq(Quat q0, Quat q1) {
  float w = qdot(q0, q1);
  if (w > 1.f) w *= rsqrtf(qdot(q0, q0) * qdot(q1, q1));
  const float s2 = 1.f - w * w;
  if (!(s2 > 1e-8f)) return 0.f;
  float theta = 2.f * fast_acos(w);
  if (theta > 3.14159265358979f) theta -= 6.28318530717959f;
  return theta * theta;
}

// Eigen 3.3.7 QuaternionBase::slerp (see slerp in imitation.cuh)
__device__ __forceinline__ Quat fast_slerp(Quat a, Quat b, float t) {
  const float d = qdot(a, b);
  const float ad = fabsf(d);
  // both branches are evaluated and selected (no divergence): the trigonometric weights for |d| < 1 - eps, the
  // linear ones otherwise (theta is clamped so that 1 / sin(theta) stays finite in the unselected lane)
  const bool lin = !(ad < 1.f - 1.1920929e-7f);
  const float theta = fast_acos(fminf(ad, 0.99999988f));
  const float inv = rcp_approx(__sinf(theta));
  const float u0 = 1.f - t;
  const float s0 = lin ? u0 : __sinf(u0 * theta) * inv;
  float s1 = lin ? t : __sinf(t * theta) * inv;
  s1 = d < 0.f ? -s1 : s1;
  return Quat{fmaf(s0, a.w, s1 * b.w), fmaf(s0, a.x, s1 * b.x), fmaf(s0, a.y, s1 * b.y), fmaf(s0, a.z, s1 * b.z)};
}

__device__ __forceinline__ Quat q4(float4 v) { return Quat{v.x, v.y, v.z, v.w}; }
__device__ __forceinline__ float lerp(float a, float b, float t) { return fmaf(t, b - a, a); }
__device__ __forceinline__ float4 lerp4(float4 a, float4 b, float t) {
  return make_float4(lerp(a.x, b.x, t), lerp(a.y, b.y, t), lerp(a.z, b.z, t), lerp(a.w, b.w, t));
}
__device__ __forceinline__ float comp(float4 v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

// cMotion::CalcIndexBlend (Motion.cpp:498-527) + CalcCycleCount (:488-496), as index_blend in imitation.cuh; the
// cycle count keeps the IEEE division (exact cycle boundaries), the blend weight uses the SFU reciprocal.
__device__ __forceinline__ void index_blend_fast(const ImitConst& c, const float* times, float time, int& idx,
                                                 float& blend, int& cycles) {
  const float dur = c.duration;
  cycles = 0;
  if (!c.loop_wrap) {
    if (time <= 0.f) { idx = 0; blend = 0.f; return; }
    if (time >= dur) { idx = c.n_frames - 2; blend = 1.f; cycles = 1; return; }
  }
  cycles = static_cast<int>(floorf(time / dur));
  if (!c.loop_wrap) cycles = min(max(cycles, 0), 1);
  time -= cycles * dur;
  int lo = 0, hi = c.n_frames;  // upper_bound(times, time) - 1
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (times[mid] <= time) lo = mid + 1; else hi = mid;
  }
  idx = min(max(lo - 1, 0), c.n_frames - 2);
  const float t0 = times[idx], t1 = times[idx + 1];
  blend = fminf(fmaxf((time - t0) * rcp_approx(t1 - t0), 0.f), 1.f);  // BlendFrames saturates (Motion.cpp:252)
}

// ---- packed pair arithmetic ------------------------------------------------------------------------------
// The simulated and the kinematic character walk the same tree with the same formulas, so their states travel as
// (sim, kin) pairs through Blackwell's packed fp32 instructions (FFMA2 / FADD2 / FMUL2: one issue slot, two
// lanes of work; negation, immediates and scalar broadcast are free source modifiers).
struct P2 { float2 v; };
__device__ __forceinline__ P2 mk(float a, float b) { return P2{make_float2(a, b)}; }
__device__ __forceinline__ P2 splat(float a) { return P2{make_float2(a, a)}; }
__device__ __forceinline__ P2 operator-(P2 a) { return P2{make_float2(-a.v.x, -a.v.y)}; }
__device__ __forceinline__ P2 operator+(P2 a, P2 b) { return P2{__fadd2_rn(a.v, b.v)}; }
__device__ __forceinline__ P2 operator-(P2 a, P2 b) { return P2{__fadd2_rn(a.v, (-b).v)}; }
__device__ __forceinline__ P2 operator*(P2 a, P2 b) { return P2{__fmul2_rn(a.v, b.v)}; }
__device__ __forceinline__ P2 operator*(P2 a, float b) { return P2{__fmul2_rn(a.v, make_float2(b, b))}; }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { return P2{__ffma2_rn(a.v, b.v, c.v)}; }
__device__ __forceinline__ P2 fma2(P2 a, float b, P2 c) { return P2{__ffma2_rn(a.v, make_float2(b, b), c.v)}; }

struct Quat2 { P2 w, x, y, z; };
struct Vec2 { P2 x, y, z; };
struct Mat2 { P2 xx, xy, xz, yx, yy, yz, zx, zy, zz; };
__device__ __forceinline__ Quat2 pack(Quat a, Quat b) { return Quat2{mk(a.w, b.w), mk(a.x, b.x), mk(a.y, b.y), mk(a.z, b.z)}; }
__device__ __forceinline__ Vec2 pack(Vec3 a, Vec3 b) { return Vec2{mk(a.x, b.x), mk(a.y, b.y), mk(a.z, b.z)}; }
__device__ __forceinline__ Vec2 operator+(Vec2 a, Vec2 b) { return Vec2{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ Vec2 cross(Vec2 a, Vec2 b) {
  return Vec2{fma2(a.y, b.z, -(a.z * b.y)), fma2(a.z, b.x, -(a.x * b.z)), fma2(a.x, b.y, -(a.y * b.x))};
}
__device__ __forceinline__ Quat2 qmul(Quat2 a, Quat2 b) {
  return Quat2{fma2(a.w, b.w, -fma2(a.x, b.x, fma2(a.y, b.y, a.z * b.z))),
               fma2(a.w, b.x, fma2(a.x, b.w, fma2(a.y, b.z, -(a.z * b.y)))),
               fma2(a.w, b.y, fma2(a.y, b.w, fma2(a.z, b.x, -(a.x * b.z)))),
               fma2(a.w, b.z, fma2(a.z, b.w, fma2(a.x, b.y, -(a.y * b.x))))};
}
// rotation matrices of a pair of UNIT quaternions
__device__ __forceinline__ Mat2 qmat(Quat2 q) {
  const P2 x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
  const P2 xx = q.x * x2, yy = q.y * y2, zz = q.z * z2;
  const P2 xy = q.x * y2, xz = q.x * z2, yz = q.y * z2;
  const P2 wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
  const P2 one = splat(1.f);
  Mat2 m;
  m.xx = one - (yy + zz); m.xy = xy - wz;         m.xz = xz + wy;
  m.yx = xy + wz;         m.yy = one - (xx + zz); m.yz = yz - wx;
  m.zx = xz - wy;         m.zy = yz + wx;         m.zz = one - (xx + yy);
  return m;
}
__device__ __forceinline__ Vec2 mul(const Mat2& m, Vec2 v) {
  return Vec2{fma2(m.xx, v.x, fma2(m.xy, v.y, m.xz * v.z)), fma2(m.yx, v.x, fma2(m.yy, v.y, m.yz * v.z)),
              fma2(m.zx, v.x, fma2(m.zy, v.y, m.zz * v.z))};
}
// m * (compile-time constant vector): zero components cost nothing
template <int J, bool BODY>
__device__ __forceinline__ Vec2 mul_const(const Mat2& m) {
  constexpr float cx = BODY ? jbody(J, 0) : jattach(J, 0);
  constexpr float cy = BODY ? jbody(J, 1) : jattach(J, 1);
  constexpr float cz = BODY ? jbody(J, 2) : jattach(J, 2);
  Vec2 r = Vec2{splat(0.f), splat(0.f), splat(0.f)};
  if constexpr (cx != 0.f) { r.x = m.xx * cx; r.y = m.yx * cx; r.z = m.zx * cx; }
  if constexpr (cy != 0.f) {
    if constexpr (cx != 0.f) { r.x = fma2(m.xy, cy, r.x); r.y = fma2(m.yy, cy, r.y); r.z = fma2(m.zy, cy, r.z); }
    else { r.x = m.xy * cy; r.y = m.yy * cy; r.z = m.zy * cy; }
  }
  if constexpr (cz != 0.f) {
    if constexpr (cx != 0.f || cy != 0.f) { r.x = fma2(m.xz, cz, r.x); r.y = fma2(m.yz, cz, r.y); r.z = fma2(m.zz, cz, r.z); }
    else { r.x = m.xz * cz; r.y = m.yz * cz; r.z = m.zz * cz; }
  }
  return r;
}

struct JointState { Quat2 Q; Vec2 p, w, v; };  // .x lanes: simulated character, .y lanes: kinematic character

// Everything one thread carries through the tree walk.  All indices are compile-time constants after
// inlining, so the arrays are promoted to registers and dead members disappear.
struct Walk {
  const float* p0;      // sim pose row (shared memory)
  const float* w0;      // sim velocity row
  const float4* f0;     // clip frame idx, permuted slots
  const float4* f1;
  const float4* g0;     // clip frame velocity idx (nullptr: zero velocity past the end of a non-looping clip)
  const float4* g1;
  float blend;
  JointState S[kJoints];
  float pose_err, vel_err, ee_err;
  Vec2 comv;
  P2 rp_x, rp_z;        // root positions (x, z); y is replaced by the height above each character's ground
  P2 ground;            // (0, kin ground height)
  P2 hc, hs;            // cos / sin of each character's heading
};

template <int J>
__device__ __forceinline__ void joint_step(Walk& s, const ImitConst& c) {
  constexpr int type = jtype(J);
  constexpr int par = jparent(J);
  constexpr int off = joff(J);
  const float jw = c.joint_w[J];
  Quat q0l = Quat{1.f, 0.f, 0.f, 0.f}, q1l = q0l;  // local joint rotations (sim / kin)
  Vec3 wl0 = v3(0, 0, 0), wl1 = v3(0, 0, 0);       // local joint angular velocities
  if constexpr (type == 1) {
    const Quat q0r = Quat{s.p0[off], s.p0[off + 1], s.p0[off + 2], s.p0[off + 3]};
    // slerp of unit quaternions is unit in exact arithmetic; the SFU sines leave a norm error of up to 1e-3 at
    // tiny frame-to-frame angles, which QuatTheta would read as a rotation -> normalise before the error term
    q1l = qnormalize(fast_slerp(q4(s.f0[jslot(J)]), q4(s.f1[jslot(J)]), s.blend));
    s.pose_err = fmaf(jw, theta_sq(q0r, q1l), s.pose_err);
    q0l = qnormalize(q0r);
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s.g0) kv = lerp4(s.g0[jslot(J)], s.g1[jslot(J)], s.blend);
    wl0 = v3(s.w0[off], s.w0[off + 1], s.w0[off + 2]);
    wl1 = v3(kv.x, kv.y, kv.z);
    const float d0 = kv.x - wl0.x, d1 = kv.y - wl0.y, d2 = kv.z - wl0.z, d3 = kv.w - s.w0[off + 3];
    s.vel_err = fmaf(jw, fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, d0 * d0))), s.vel_err);
  } else if constexpr (type == 2) {
    const float t0 = s.p0[off];
    const float t1 = lerp(comp(s.f0[jslot(J)], jsub(J)), comp(s.f1[jslot(J)], jsub(J)), s.blend);
    const float dt = t1 - t0;
    s.pose_err = fmaf(jw, dt * dt, s.pose_err);
    float sn, cs;
    __sincosf(0.5f * t0, &sn, &cs);
    q0l = Quat{cs, 0.f, 0.f, sn};
    __sincosf(0.5f * t1, &sn, &cs);
    q1l = Quat{cs, 0.f, 0.f, sn};
    const float kv = s.g0 ? lerp(comp(s.g0[jslot(J)], jsub(J)), comp(s.g1[jslot(J)], jsub(J)), s.blend) : 0.f;
    const float dv = kv - s.w0[off];
    s.vel_err = fmaf(jw, dv * dv, s.vel_err);
    wl0 = v3(0.f, 0.f, s.w0[off]);
    wl1 = v3(0.f, 0.f, kv);
  }
  JointState& a = s.S[J];
  const JointState& pa = s.S[par];
  // the parent's matrices are recomputed here; after inlining the compiler shares them between siblings
  const Mat2 Rp = qmat(pa.Q);
  const Vec2 r = mul_const<J, false>(Rp);
  a.p = pa.p + r;
  a.v = pa.v + cross(pa.w, r);
  if constexpr (type == 3) {
    a.Q = pa.Q;
    a.w = pa.w;
  } else if constexpr (type == 2) {
    // Q_parent * (c, 0, 0, s)
    const P2 cw = mk(q0l.w, q1l.w), sz = mk(q0l.z, q1l.z);
    a.Q = Quat2{fma2(pa.Q.w, cw, -(pa.Q.z * sz)), fma2(pa.Q.x, cw, pa.Q.y * sz), fma2(pa.Q.y, cw, -(pa.Q.x * sz)),
                fma2(pa.Q.z, cw, pa.Q.w * sz)};
    // the joint axis is local z, unchanged by the joint's own rotation: column z of the parent's matrix
    const P2 wz = mk(wl0.z, wl1.z);
    a.w = Vec2{fma2(wz, Rp.xz, pa.w.x), fma2(wz, Rp.yz, pa.w.y), fma2(wz, Rp.zz, pa.w.z)};
  } else {
    a.Q = qmul(pa.Q, pack(q0l, q1l));
  }
  constexpr bool need_mat = (type == 1) || jbody(J, 0) != 0.f || jbody(J, 1) != 0.f || jbody(J, 2) != 0.f;
  const float mf = c.mass_frac[J];
  if constexpr (need_mat) {
    const Mat2 R = qmat(a.Q);
    if constexpr (type == 1) a.w = pa.w + mul(R, pack(wl0, wl1));
    // centre-of-mass velocity (RBDUtil.cpp:572-613)
    const Vec2 cv = a.v + cross(a.w, mul_const<J, true>(R));
    s.comv = Vec2{fma2(cv.x, mf, s.comv.x), fma2(cv.y, mf, s.comv.y), fma2(cv.z, mf, s.comv.z)};
  } else {
    s.comv = Vec2{fma2(a.v.x, mf, s.comv.x), fma2(a.v.y, mf, s.comv.y), fma2(a.v.z, mf, s.comv.z)};
  }
  if constexpr (jee(J)) {  // SceneImitate.cpp:78-96
    // position relative to the root, height above the character's own ground, rotated by -heading about y
    const P2 rx = a.p.x - s.rp_x, ry = a.p.y - s.ground, rz = a.p.z - s.rp_z;
    const P2 hx = fma2(s.hc, rx, -(s.hs * rz)), hz = fma2(s.hs, rx, s.hc * rz);
    const float dx = hx.v.y - hx.v.x, dy = ry.v.y - ry.v.x, dz = hz.v.y - hz.v.x;
    s.ee_err += fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  }
}

// Tree-walk order: any topological order gives the same sums; this one finishes the chains that hang off the
// hips first, so that only one branch-point state (root, then chest) plus the current chain is live at any
// time (keeps the kernel inside 128 registers).
template <int... Js>
__device__ __forceinline__ void walk_joints(Walk& s, const ImitConst& c, std::integer_sequence<int, Js...>) {
  (joint_step<Js>(s, c), ...);
}
using WalkOrder = std::integer_sequence<int, 3, 4, 5, 9, 10, 11, 1, 2, 6, 7, 8, 12, 13, 14>;

constexpr int kTimesPad = 128;      // smem floats reserved for frame times
constexpr int kWarps = 16;          // one block per SM; every warp runs its own staging pipeline
constexpr int kBlock = kWarps * 32;
constexpr int kTileRows = 32;       // rows per warp tile: 32 x 172 B = 5504 B, a multiple of 16
constexpr int kTileFloats = kTileRows * kDof;

inline size_t smem_bytes(int n_frames) {
  return size_t(kTimesPad) * 4 + 2 * size_t(n_frames) * kSlots * 16 + size_t(kWarps) * 2 * kTileFloats * 4 +
         size_t(kWarps) * 8;
}

__global__ void __launch_bounds__(kBlock, 1)
imitation_reward_h3d_kernel(const __grid_constant__ ImitConst c, const float* __restrict__ g_times,
                            const float4* __restrict__ g_frames, const float4* __restrict__ g_fvel,
                            const float* __restrict__ pose, const float* __restrict__ vel,
                            const float* __restrict__ kin_time, const float* __restrict__ kin_origin,
                            long long n_envs, float* __restrict__ reward, float* __restrict__ terms, int bulk_ok) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* s_times = reinterpret_cast<float*>(sm_raw);
  float4* s_frames = reinterpret_cast<float4*>(s_times + kTimesPad);
  float4* s_fvel = s_frames + c.n_frames * kSlots;
  float* s_rows = reinterpret_cast<float*>(s_fvel + c.n_frames * kSlots);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rows + kWarps * 2 * kTileFloats);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < c.n_frames; i += kBlock) s_times[i] = g_times[i];
  for (int i = threadIdx.x; i < c.n_frames * kSlots; i += kBlock) {
    s_frames[i] = g_frames[i];
    s_fvel[i] = g_fvel[i];
  }
  if (lane == 0) ptx::mbar_init(&bars[warp], 1);
  if (threadIdx.x == 0) ptx::fence_barrier_init();
  __syncthreads();  // the only block-wide barrier: clip tables and mbarriers are ready

  float* s_pose = s_rows + warp * 2 * kTileFloats;
  float* s_vel = s_pose + kTileFloats;
  uint64_t* bar = &bars[warp];
  uint32_t parity = 0;
  const long long n_tiles = (n_envs + kTileRows - 1) / kTileRows;
  const long long warp_stride = static_cast<long long>(gridDim.x) * kWarps;

  for (long long tile = static_cast<long long>(blockIdx.x) * kWarps + warp; tile < n_tiles; tile += warp_stride) {
    const long long base = tile * kTileRows;
    const int rows = static_cast<int>(min(static_cast<long long>(kTileRows), n_envs - base));
    const uint32_t bytes = static_cast<uint32_t>(rows) * kDof * 4;
    const bool bulk = bulk_ok && (bytes & 15u) == 0;
    __syncwarp();  // every lane is done with the previous tile's rows
    if (bulk) {
      if (lane == 0) {
        ptx::fence_proxy_async_smem();  // order the previous tile's generic reads before the async writes
        ptx::mbar_arrive_expect_tx(bar, 2 * bytes);
        ptx::bulk_load_1d(s_pose, pose + base * kDof, bytes, bar);
        ptx::bulk_load_1d(s_vel, vel + base * kDof, bytes, bar);
        // pull this warp's next tile into L2 while the current one is processed (single smem stage per warp)
        const long long nbase = base + warp_stride * kTileRows;
        if (nbase + kTileRows <= n_envs) {
          ptx::bulk_prefetch_l2(pose + nbase * kDof, kTileFloats * 4);
          ptx::bulk_prefetch_l2(vel + nbase * kDof, kTileFloats * 4);
        }
      }
    } else {
      const float* gp = pose + base * kDof;
      const float* gv = vel + base * kDof;
      for (int i = lane; i < rows * kDof; i += 32) {
        s_pose[i] = gp[i];
        s_vel[i] = gv[i];
      }
    }
    const long long e = base + lane;
    // per-env scalars are fetched while the rows are in flight
    const bool live = lane < rows;
    const float time_e = live ? kin_time[e] : 0.f;
    Vec3 org = v3(0.f, 0.f, 0.f);
    if (live && kin_origin) org = v3(kin_origin[e * 3], kin_origin[e * 3 + 1], kin_origin[e * 3 + 2]);
    if (bulk) {
      ptx::mbar_wait(bar, parity);
      parity ^= 1;
    } else {
      __syncwarp();
    }
    if (!live) continue;

    Walk s;
    s.p0 = s_pose + lane * kDof;
    s.w0 = s_vel + lane * kDof;
    int idx, cycles;
    index_blend_fast(c, s_times, time_e, idx, s.blend, cycles);
    s.f0 = s_frames + idx * kSlots;
    s.f1 = s.f0 + kSlots;
    s.g0 = s_fvel + idx * kSlots;
    s.g1 = s.g0 + kSlots;
    if (!c.loop_wrap && time_e >= c.duration) s.g0 = s.g1 = nullptr;  // Motion.cpp:278-281
    const float cyc = c.loop_wrap ? static_cast<float>(cycles) : 0.f;     // MotionController.cpp:29-40
    const Vec3 root_off = v3(fmaf(cyc, c.cycle_delta[0], org.x), fmaf(cyc, c.cycle_delta[1], org.y),
                             fmaf(cyc, c.cycle_delta[2], org.z));
    const float ground_h1 = org.y;  // kin_char.GetOriginPos()[1]; the sim ground is the plane y = 0
    s.pose_err = s.vel_err = s.ee_err = 0.f;

    // ---- root (ChildParentTransRoot, BuildJointSubspaceRoot) ----
    const Quat rq0_raw = Quat{s.p0[3], s.p0[4], s.p0[5], s.p0[6]};
    Quat rq1 = qnormalize(fast_slerp(q4(s.f0[1]), q4(s.f1[1]), s.blend));  // KinTree.cpp:1595-1596
    if (rq1.w < 0.f) rq1 = Quat{-rq1.w, -rq1.x, -rq1.y, -rq1.z};            // StandardizeQuat, KinCharacter.cpp:585
    const Vec3 rp0 = v3(s.p0[0], s.p0[1], s.p0[2]);
    const float4 kp = lerp4(s.f0[0], s.f1[0], s.blend);
    const Vec3 rp1 = v3(kp.x, kp.y, kp.z) + root_off;
    float4 kv_lin = make_float4(0.f, 0.f, 0.f, 0.f), kv_ang = kv_lin;
    if (s.g0) {
      kv_lin = lerp4(s.g0[0], s.g1[0], s.blend);
      kv_ang = lerp4(s.g0[1], s.g1[1], s.blend);
    }
    const Vec3 v0 = v3(s.w0[0], s.w0[1], s.w0[2]), v1 = v3(kv_lin.x, kv_lin.y, kv_lin.z);
    s.S[0].Q = pack(qnormalize(rq0_raw), rq1);
    s.S[0].p = pack(rp0, rp1);
    s.S[0].v = pack(v0, v1);
    s.S[0].w = pack(v3(s.w0[3], s.w0[4], s.w0[5]), v3(kv_ang.x, kv_ang.y, kv_ang.z));
    s.rp_x = s.S[0].p.x;
    s.rp_z = s.S[0].p.z;
    s.ground = mk(0.f, ground_h1);
    float hc0, hs0, hc1, hs1;
    heading_cs(rq0_raw, hc0, hs0);
    heading_cs(rq1, hc1, hs1);
    s.hc = mk(hc0, hc1);
    s.hs = mk(hs0, hs1);
    const float root_rot_err = theta_sq(rq0_raw, rq1);
    float root_ang_err;
    {
      const float d0 = kv_ang.x - s.w0[3], d1 = kv_ang.y - s.w0[4], d2 = kv_ang.z - s.w0[5], d3 = kv_ang.w - s.w0[6];
      root_ang_err = fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, d0 * d0)));
    }
    s.pose_err = c.joint_w[0] * root_rot_err;       // SceneImitate.cpp:66-68
    s.vel_err = c.joint_w[0] * root_ang_err;
    {
      const Mat2 R = qmat(s.S[0].Q);
      const Vec2 cv = s.S[0].v + cross(s.S[0].w, mul_const<0, true>(R));
      s.comv = Vec2{cv.x * c.mass_frac[0], cv.y * c.mass_frac[0], cv.z * c.mass_frac[0]};
    }
    // root terms (SceneImitate.cpp:99-114)
    const Vec3 dr = rp0 - v3(rp1.x, rp1.y - ground_h1, rp1.z);
    const Vec3 dv = v1 - v0;
    const float root_err = dot(dr, dr) + 0.1f * root_rot_err + 0.01f * dot(dv, dv) + 0.001f * root_ang_err;

    walk_joints(s, c, WalkOrder{});

    const Vec3 dc = v3(s.comv.x.v.y - s.comv.x.v.x, s.comv.y.v.y - s.comv.y.v.x, s.comv.z.v.y - s.comv.z.v.x);
    const float com_err = 0.1f * dot(dc, dc);  // SceneImitate.cpp:115

    const float r_pose = __expf(-c.s_pose * s.pose_err);
    const float r_vel = __expf(-c.s_vel * s.vel_err);
    const float r_ee = __expf(-c.s_ee * s.ee_err);
    const float r_root = __expf(-c.s_root * root_err);
    const float r_com = __expf(-c.s_com * com_err);
    reward[e] = c.w_pose * r_pose + c.w_vel * r_vel + c.w_ee * r_ee + c.w_root * r_root + c.w_com * r_com;
    if (terms) {
      float* t = terms + e * 5;
      t[0] = r_pose; t[1] = r_vel; t[2] = r_ee; t[3] = r_root; t[4] = r_com;
    }
  }
}

}  // namespace h3d
}  // namespace simstep
