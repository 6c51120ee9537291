// DeepMimic motion-imitation reward, batched: one thread per env.
//
// Follows reference DeepMimicCore scenes/SceneImitate.cpp:7-127 (CalcRewardImitate) with the kinematic
// chain of anim/KinTree.cpp (ChildParentTrans* :1806-1878, JointWorldTrans :1126-1139, CalcPoseErr
// :1367-1398, CalcVelErr :1445-1462, CalcHeading :1667-1675), util/MathUtil.cpp (QuatTheta :538-554,
// RotateMat(q) :212-242), the centre-of-mass velocity of sim/RBDUtil.cpp:572-613 and the clip sampling
// of anim/Motion.cpp:267-305, :498-527 + anim/KinTree.cpp:1577-1625 (slerp as in Eigen 3.3.7) +
// anim/MotionController.cpp:25-41 + anim/KinCharacter.cpp:573-640.
//
// The 4x4 / 6-D spatial algebra of the reference is evaluated in its closed form for rigid trees with
// zero attach rotations:
//   Q_j = Q_parent * q_j,  p_j = p_parent + Q_parent.attach_j,
//   w_j = w_parent + Q_j.w_local_j,  v_j = v_parent + w_parent x (p_j - p_parent),
//   com_j = p_j + Q_j.body_attach_j,  d/dt com_j = v_j + w_j x (com_j - p_j).
// Both characters (simulated from the inputs, kinematic from the clip) walk the tree in one loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/simstep.h"

namespace simstep {

struct ImitConst {
  int n_joints;
  int dof;
  int n_frames;
  int loop_wrap;
  float duration;
  float cycle_delta[3];
  int joint_type[SIMSTEP_MAX_JOINTS];
  int parent[SIMSTEP_MAX_JOINTS];
  int param_offset[SIMSTEP_MAX_JOINTS];
  int is_end_eff[SIMSTEP_MAX_JOINTS];
  float attach[SIMSTEP_MAX_JOINTS][3];
  float body_attach[SIMSTEP_MAX_JOINTS][3];
  float mass_frac[SIMSTEP_MAX_JOINTS];  // body mass / total mass
  float joint_w[SIMSTEP_MAX_JOINTS];    // DiffWeight / sum|DiffWeight| (SceneImitate.cpp:300-312)
  float w_pose, w_vel, w_ee, w_root, w_com;                    // normalised reward weights (:9-20)
  float s_pose, s_vel, s_ee, s_root, s_com;                    // exponent scales (:25-30)
};

constexpr int kImitThreads = 128;

struct Quat { float w, x, y, z; };
struct Vec3 { float x, y, z; };

__device__ __forceinline__ Vec3 v3(float x, float y, float z) { return Vec3{x, y, z}; }
__device__ __forceinline__ Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ Vec3 operator*(float s, Vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float qdot(Quat a, Quat b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Quat qmul(Quat a, Quat b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
// rotation of v by a UNIT quaternion
__device__ __forceinline__ Vec3 qrot(Quat q, Vec3 v) {
  const Vec3 u = v3(q.x, q.y, q.z);
  const Vec3 t = 2.f * cross(u, v);
  return v + q.w * t + cross(u, t);
}
// RotateMat(q) divides by |q|^2 (MathUtil.cpp:220): equivalent to normalising q first
__device__ __forceinline__ Quat qnormalize(Quat q) {
  const float inv = rsqrtf(qdot(q, q));
  return Quat{q.w * inv, q.x * inv, q.y * inv, q.z * inv};
}

// QuatTheta(QuatDiff(q0, q1))^2 (MathUtil.cpp:527-554, :33-46): the w component of q1*conj(q0) is <q0,q1>
__device__ __forceinline__ float quat_theta_sq(Quat q0, Quat q1) {
  float w = qdot(q0, q1);
  if (w > 1.f) w *= rsqrtf(qdot(q0, q0) * qdot(q1, q1));
  const float s2 = 1.f - w * w;
  if (!(s2 > 1e-8f)) return 0.f;  // sin_theta <= 1e-4 (or NaN): theta = 0
  float theta = 2.f * acosf(w);
  if (theta > 3.14159265358979f) theta -= 6.28318530717959f;
  return theta * theta;
}

// Eigen 3.3.7 QuaternionBase::slerp
__device__ __forceinline__ Quat slerp(Quat a, Quat b, float t) {
  const float d = qdot(a, b);
  const float ad = fabsf(d);
  float s0, s1;
  if (ad >= 1.f - 1.1920929e-7f) {
    s0 = 1.f - t;
    s1 = t;
  } else {
    const float theta = acosf(ad);
    const float inv = 1.f / sinf(theta);
    s0 = sinf((1.f - t) * theta) * inv;
    s1 = sinf(t * theta) * inv;
  }
  if (d < 0.f) s1 = -s1;
  return Quat{s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z};
}

__device__ __forceinline__ Quat load_quat(const float* p) { return Quat{p[0], p[1], p[2], p[3]}; }

// cMotion::CalcIndexBlend (Motion.cpp:498-527) + CalcCycleCount (:488-496)
__device__ __forceinline__ void index_blend(const ImitConst& c, const float* times, float time, int& idx, float& blend,
                                            int& cycles) {
  const float dur = c.duration;
  cycles = 0;
  if (!c.loop_wrap) {
    if (time <= 0.f) { idx = 0; blend = 0.f; return; }
    if (time >= dur) { idx = c.n_frames - 2; blend = 1.f; cycles = 1; return; }
  }
  cycles = static_cast<int>(floorf(time / dur));
  if (!c.loop_wrap) cycles = min(max(cycles, 0), 1);
  time -= cycles * dur;
  // upper_bound(times, time) - 1
  int lo = 0, hi = c.n_frames;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (times[mid] <= time) lo = mid + 1; else hi = mid;
  }
  idx = min(max(lo - 1, 0), c.n_frames - 2);
  const float t0 = times[idx], t1 = times[idx + 1];
  blend = (time - t0) / (t1 - t0);
  blend = fminf(fmaxf(blend, 0.f), 1.f);  // BlendFrames saturates (Motion.cpp:252)
}

struct JointState { Quat Q; Vec3 p, w, v; };

// Shared-memory layout of one block: clip tables, then the block's sim pose/vel rows (coalesced load).
struct ImitSmem {
  float* times;   // [n_frames]
  float* frames;  // [n_frames][dof]
  float* fvel;    // [n_frames][dof]
  float* pose;    // [kImitThreads][dof]
  float* vel;     // [kImitThreads][dof]
};

__device__ __forceinline__ ImitSmem carve_smem(float* base, const ImitConst& c, bool with_rows) {
  ImitSmem s;
  s.times = base;
  s.frames = s.times + ((c.n_frames + 3) & ~3);
  s.fvel = s.frames + c.n_frames * c.dof;
  s.pose = s.fvel + c.n_frames * c.dof;
  s.vel = with_rows ? s.pose + kImitThreads * c.dof : s.pose;
  return s;
}

inline size_t imit_smem_bytes(int n_frames, int dof, bool with_rows) {
  size_t f = ((n_frames + 3) & ~3) + 2 * size_t(n_frames) * dof;
  if (with_rows) f += 2 * size_t(kImitThreads) * dof;
  return f * sizeof(float);
}

__device__ __forceinline__ void load_clip_to_smem(const ImitSmem& s, const ImitConst& c, const float* g_times,
                                                  const float* g_frames, const float* g_fvel) {
  for (int i = threadIdx.x; i < c.n_frames; i += blockDim.x) s.times[i] = g_times[i];
  const int n = c.n_frames * c.dof;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s.frames[i] = g_frames[i];
    s.fvel[i] = g_fvel[i];
  }
}

// Kinematic character's local joint rotation / velocity parameters at a clip time.
struct KinSampler {
  const float* f0;  // frame idx
  const float* f1;  // frame idx+1
  const float* v0;
  const float* v1;
  float blend;
  Vec3 root_off;    // cycle offset + kin origin

  __device__ __forceinline__ Vec3 root_pos() const {
    return v3((1.f - blend) * f0[0] + blend * f1[0], (1.f - blend) * f0[1] + blend * f1[1],
              (1.f - blend) * f0[2] + blend * f1[2]) + root_off;
  }
  __device__ __forceinline__ Quat root_rot() const {
    Quat q = qnormalize(slerp(load_quat(f0 + 3), load_quat(f1 + 3), blend));  // KinTree.cpp:1595-1596
    if (q.w < 0.f) q = Quat{-q.w, -q.x, -q.y, -q.z};                       // StandardizeQuat, KinCharacter.cpp:585
    return q;
  }
  __device__ __forceinline__ Quat joint_rot(int off) const { return slerp(load_quat(f0 + off), load_quat(f1 + off), blend); }
  __device__ __forceinline__ float scalar(int off) const { return (1.f - blend) * f0[off] + blend * f1[off]; }
  __device__ __forceinline__ float vel(int off) const { return (1.f - blend) * v0[off] + blend * v1[off]; }
};

__device__ __forceinline__ KinSampler make_sampler(const ImitConst& c, const ImitSmem& s, float time,
                                                   const float* origin /* 3 or null */) {
  int idx, cycles;
  float blend;
  index_blend(c, s.times, time, idx, blend, cycles);
  KinSampler k;
  k.f0 = s.frames + idx * c.dof;
  k.f1 = k.f0 + c.dof;
  k.v0 = s.fvel + idx * c.dof;
  k.v1 = k.v0 + c.dof;
  k.blend = blend;
  const float cyc = c.loop_wrap ? static_cast<float>(cycles) : 0.f;  // MotionController.cpp:29-40
  k.root_off = v3(cyc * c.cycle_delta[0], cyc * c.cycle_delta[1], cyc * c.cycle_delta[2]);
  if (origin) k.root_off = k.root_off + v3(origin[0], origin[1], origin[2]);
  // CalcFrameVel zeroes the velocity past the end of a non-looping clip (Motion.cpp:278-281)
  if (!c.loop_wrap && time >= c.duration) { k.v0 = k.v1 = nullptr; }
  return k;
}

// heading frame: rotation by -heading about y (KinTree.cpp:1667-1712), applied to a direction.  The reference
// rotates the x axis with Eigen's q*v, which does NOT normalise q: pass the raw root quaternion.
__device__ __forceinline__ void heading_cs(Quat root_q, float& hc, float& hs) {
  const Vec3 d = qrot(root_q, v3(1.f, 0.f, 0.f));
  const float n2 = d.x * d.x + d.z * d.z;
  if (n2 > 0.f) {
    const float inv = rsqrtf(n2);
    hc = d.x * inv;   // cos(heading)
    hs = -d.z * inv;  // sin(heading)
  } else {
    hc = 1.f;
    hs = 0.f;  // atan2(0, 0) = 0
  }
}
__device__ __forceinline__ Vec3 rot_heading_inv(float hc, float hs, Vec3 v) {
  // RotateMat(y, -heading): x' = c x - s z, z' = s x + c z with c = cos(h), s = sin(h)
  return v3(hc * v.x - hs * v.z, v.y, hs * v.x + hc * v.z);
}

__global__ void __launch_bounds__(kImitThreads)
imitation_reward_kernel(const ImitConst c, const float* __restrict__ g_times, const float* __restrict__ g_frames,
                        const float* __restrict__ g_fvel, const float* __restrict__ pose, const float* __restrict__ vel,
                        const float* __restrict__ kin_time, const float* __restrict__ kin_origin, long long n_envs,
                        float* __restrict__ reward, float* __restrict__ terms) {
  extern __shared__ float sm_f[];
  const ImitSmem s = carve_smem(sm_f, c, true);
  load_clip_to_smem(s, c, g_times, g_frames, g_fvel);
  const int dof = c.dof;

  for (long long base = static_cast<long long>(blockIdx.x) * kImitThreads; base < n_envs;
       base += static_cast<long long>(gridDim.x) * kImitThreads) {
    __syncthreads();  // clip tables ready / previous rows consumed
    const long long rows = min(static_cast<long long>(kImitThreads), n_envs - base);
    const int n = static_cast<int>(rows) * dof;
    const float* gp = pose + base * dof;
    const float* gv = vel + base * dof;
    for (int i = threadIdx.x; i < n; i += kImitThreads) {
      s.pose[i] = gp[i];
      s.vel[i] = gv[i];
    }
    __syncthreads();
    const long long e = base + threadIdx.x;
    if (e >= n_envs) continue;

    const float* p0 = s.pose + threadIdx.x * dof;
    const float* w0 = s.vel + threadIdx.x * dof;
    const float* org = kin_origin ? kin_origin + e * 3 : nullptr;
    const KinSampler k = make_sampler(c, s, kin_time[e], org);
    const float ground_h1 = org ? org[1] : 0.f;  // kin_char.GetOriginPos()[1]; the sim ground is the plane y = 0

    JointState A[SIMSTEP_MAX_JOINTS];  // simulated character
    JointState B[SIMSTEP_MAX_JOINTS];  // kinematic character
    float pose_err = 0.f, vel_err = 0.f, ee_err = 0.f;
    Vec3 comv0 = v3(0, 0, 0), comv1 = v3(0, 0, 0);
    float hc0, hs0, hc1, hs1;
    Quat rq0_raw, rq1;
    Vec3 rp0, rp1;

#pragma unroll 1
    for (int j = 0; j < c.n_joints; ++j) {
      const int type = c.joint_type[j];
      const int off = c.param_offset[j];
      const float jw = c.joint_w[j];
      Quat q0l, q1l;   // local joint rotations (sim / kin)
      Vec3 wl0 = v3(0, 0, 0), wl1 = v3(0, 0, 0);
      if (type == 0) {  // root: ChildParentTransRoot, BuildJointSubspaceRoot
        rq0_raw = load_quat(p0 + 3);
        rq1 = k.root_rot();
        rp0 = v3(p0[0], p0[1], p0[2]);
        rp1 = k.root_pos();
        A[0].Q = qnormalize(rq0_raw);
        B[0].Q = rq1;
        A[0].p = rp0;
        B[0].p = rp1;
        A[0].v = v3(w0[0], w0[1], w0[2]);
        A[0].w = v3(w0[3], w0[4], w0[5]);
        float kv[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) kv[i] = k.v0 ? k.vel(i) : 0.f;
        B[0].v = v3(kv[0], kv[1], kv[2]);
        B[0].w = v3(kv[3], kv[4], kv[5]);
        heading_cs(rq0_raw, hc0, hs0);
        heading_cs(B[0].Q, hc1, hs1);
        // SceneImitate.cpp:66-68: root rotation and root angular velocity (4 stored components)
        pose_err += jw * quat_theta_sq(rq0_raw, rq1);
        float dv = 0.f;
#pragma unroll
        for (int i = 3; i < 7; ++i) dv += (kv[i] - w0[i]) * (kv[i] - w0[i]);
        vel_err += jw * dv;
      } else {
        const int par = c.parent[j];
        const Vec3 at = v3(c.attach[j][0], c.attach[j][1], c.attach[j][2]);
        float perr = 0.f, verr = 0.f;
        if (type == 1) {  // spherical
          const Quat q0r = load_quat(p0 + off);
          const Quat q1r = k.joint_rot(off);
          perr = quat_theta_sq(q0r, q1r);
          q0l = qnormalize(q0r);
          q1l = qnormalize(q1r);
          float kv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            kv[i] = k.v0 ? k.vel(off + i) : 0.f;
            const float d = kv[i] - w0[off + i];
            verr += d * d;
          }
          wl0 = v3(w0[off], w0[off + 1], w0[off + 2]);
          wl1 = v3(kv[0], kv[1], kv[2]);
        } else if (type == 2) {  // revolute about local z
          const float t0 = p0[off], t1 = k.scalar(off);
          perr = (t1 - t0) * (t1 - t0);
          float s_, c_;
          sincosf(0.5f * t0, &s_, &c_);
          q0l = Quat{c_, 0.f, 0.f, s_};
          sincosf(0.5f * t1, &s_, &c_);
          q1l = Quat{c_, 0.f, 0.f, s_};
          const float kv = k.v0 ? k.vel(off) : 0.f;
          verr = (kv - w0[off]) * (kv - w0[off]);
          wl0 = v3(0.f, 0.f, w0[off]);
          wl1 = v3(0.f, 0.f, kv);
        } else {  // fixed
          q0l = Quat{1.f, 0.f, 0.f, 0.f};
          q1l = q0l;
        }
        pose_err += jw * perr;
        vel_err += jw * verr;
        A[j].Q = qmul(A[par].Q, q0l);
        B[j].Q = qmul(B[par].Q, q1l);
        const Vec3 r0 = qrot(A[par].Q, at), r1 = qrot(B[par].Q, at);
        A[j].p = A[par].p + r0;
        B[j].p = B[par].p + r1;
        A[j].w = A[par].w + qrot(A[j].Q, wl0);
        B[j].w = B[par].w + qrot(B[j].Q, wl1);
        A[j].v = A[par].v + cross(A[par].w, r0);
        B[j].v = B[par].v + cross(B[par].w, r1);
      }
      // centre-of-mass velocity (RBDUtil.cpp:572-613)
      const float mf = c.mass_frac[j];
      if (mf > 0.f) {
        const Vec3 ba = v3(c.body_attach[j][0], c.body_attach[j][1], c.body_attach[j][2]);
        comv0 = comv0 + mf * (A[j].v + cross(A[j].w, qrot(A[j].Q, ba)));
        comv1 = comv1 + mf * (B[j].v + cross(B[j].w, qrot(B[j].Q, ba)));
      }
      // end effectors (SceneImitate.cpp:78-96)
      if (type != 0 && c.is_end_eff[j]) {
        Vec3 rel0 = A[j].p - rp0, rel1 = B[j].p - rp1;
        rel0.y = A[j].p.y;              // ground height under the sim character is 0 (plane)
        rel1.y = B[j].p.y - ground_h1;
        const Vec3 d = rot_heading_inv(hc1, hs1, rel1) - rot_heading_inv(hc0, hs0, rel0);
        ee_err += dot(d, d);
      }
    }

    // root and centre-of-mass terms (SceneImitate.cpp:99-115)
    Vec3 dr = rp0 - v3(rp1.x, rp1.y - ground_h1, rp1.z);
    const float root_pos_err = dot(dr, dr);
    const float root_rot_err = quat_theta_sq(rq0_raw, rq1);
    const Vec3 dv = B[0].v - A[0].v;
    float root_ang_err = 0.f;
#pragma unroll
    for (int i = 3; i < 7; ++i) {
      const float d = (k.v0 ? k.vel(i) : 0.f) - w0[i];
      root_ang_err += d * d;
    }
    const float root_err = root_pos_err + 0.1f * root_rot_err + 0.01f * dot(dv, dv) + 0.001f * root_ang_err;
    const Vec3 dc = comv1 - comv0;
    const float com_err = 0.1f * dot(dc, dc);

    const float r_pose = expf(-c.s_pose * pose_err);
    const float r_vel = expf(-c.s_vel * vel_err);
    const float r_ee = expf(-c.s_ee * ee_err);
    const float r_root = expf(-c.s_root * root_err);
    const float r_com = expf(-c.s_com * com_err);
    reward[e] = c.w_pose * r_pose + c.w_vel * r_vel + c.w_ee * r_ee + c.w_root * r_root + c.w_com * r_com;
    if (terms) {
      float* t = terms + e * 5;
      t[0] = r_pose; t[1] = r_vel; t[2] = r_ee; t[3] = r_root; t[4] = r_com;
    }
  }
}

// cKinCharacter::CalcPose / CalcVel at a clip time (KinCharacter.cpp:573-640), origin rotation = identity.
__global__ void __launch_bounds__(kImitThreads)
clip_sample_kernel(const ImitConst c, const float* __restrict__ g_times, const float* __restrict__ g_frames,
                   const float* __restrict__ g_fvel, const float* __restrict__ kin_time,
                   const float* __restrict__ kin_origin, long long n_envs, float* __restrict__ out_pose,
                   float* __restrict__ out_vel) {
  extern __shared__ float sm_f[];
  const ImitSmem s = carve_smem(sm_f, c, false);
  load_clip_to_smem(s, c, g_times, g_frames, g_fvel);
  __syncthreads();
  for (long long e = static_cast<long long>(blockIdx.x) * kImitThreads + threadIdx.x; e < n_envs;
       e += static_cast<long long>(gridDim.x) * kImitThreads) {
    const float* org = kin_origin ? kin_origin + e * 3 : nullptr;
    const KinSampler k = make_sampler(c, s, kin_time[e], org);
    float* po = out_pose + e * c.dof;
    float* vo = out_vel + e * c.dof;
    for (int j = 0; j < c.n_joints; ++j) {
      const int type = c.joint_type[j];
      const int off = c.param_offset[j];
      if (type == 0) {
        const Vec3 rp = k.root_pos();
        const Quat rq = k.root_rot();
        po[0] = rp.x; po[1] = rp.y; po[2] = rp.z;
        po[3] = rq.w; po[4] = rq.x; po[5] = rq.y; po[6] = rq.z;
        for (int i = 0; i < 7; ++i) vo[i] = k.v0 ? k.vel(i) : 0.f;
      } else if (type == 1) {
        const Quat q = k.joint_rot(off);
        po[off] = q.w; po[off + 1] = q.x; po[off + 2] = q.y; po[off + 3] = q.z;
        for (int i = 0; i < 4; ++i) vo[off + i] = k.v0 ? k.vel(off + i) : 0.f;
      } else if (type == 2) {
        po[off] = k.scalar(off);
        vo[off] = k.v0 ? k.vel(off) : 0.f;
      }
    }
  }
}

// Character state features of the learned-dynamics env (reference DeepMimicCore sim/CtController.cpp:378-495,
// cCtController::BuildStatePose / BuildStateVel, the vector SimEnv.reset() obtains through record_state,
// gym-simenv/gym_simenv/envs/sim_env.py:270-285) from a generalized pose / velocity: root height, then per body
// part its centre position, rotation normal (q * y) and tangent (q * x), then per body part its linear and angular
// velocity — in the character's heading frame (cKinTree::BuildOriginTrans, KinTree.cpp:1701-1712) unless the
// controller's Record* flags say otherwise.  Body transforms follow the pose kinematically
// (cKinTree::BodyWorldTrans, KinTree.cpp:1148-1166, zero attach rotations): what the simulator holds right after
// cSimCharacter::SetPose / SetVel, i.e. at reset.  The ground is the plane y = 0.
//   out [n][1 + 9 * n_joints + 6 * n_joints]
struct RecordFlags {
  int all_world;        // RecordAllWorld
  int world_root_pos;   // RecordWorldRootPos
  int world_root_rot;   // RecordWorldRootRot
  float vel_scale;      // 1, or 1 / UpdateRate when RecordVelAsPos
};

__global__ void __launch_bounds__(kImitThreads)
record_state_kernel(const ImitConst c, const float* __restrict__ pose, const float* __restrict__ vel,
                    long long n_envs, const RecordFlags f, float* __restrict__ out) {
  const int nj = c.n_joints;
  const int width = 1 + 15 * nj;
  const int vel0 = 1 + 9 * nj;
  for (long long e = static_cast<long long>(blockIdx.x) * kImitThreads + threadIdx.x; e < n_envs;
       e += static_cast<long long>(gridDim.x) * kImitThreads) {
    const float* p0 = pose + e * c.dof;
    const float* w0 = vel + e * c.dof;
    float* o = out + e * width;
    JointState A[SIMSTEP_MAX_JOINTS];
    const Quat rq_raw = load_quat(p0 + 3);
    A[0].Q = qnormalize(rq_raw);
    A[0].p = v3(p0[0], p0[1], p0[2]);
    A[0].v = v3(w0[0], w0[1], w0[2]);
    A[0].w = v3(w0[3], w0[4], w0[5]);
    float hc, hs;
    heading_cs(rq_raw, hc, hs);
    const Vec3 root = A[0].p;
    o[0] = root.y;  // origin_trans * (root - ground) keeps y (CtController.cpp:388-394)
#pragma unroll 1
    for (int j = 0; j < nj; ++j) {
      if (j > 0) {
        const int type = c.joint_type[j];
        const int off = c.param_offset[j];
        const int par = c.parent[j];
        Quat ql = Quat{1.f, 0.f, 0.f, 0.f};
        Vec3 wl = v3(0.f, 0.f, 0.f);
        if (type == 1) {
          ql = qnormalize(load_quat(p0 + off));
          wl = v3(w0[off], w0[off + 1], w0[off + 2]);
        } else if (type == 2) {
          float s_, c_;
          sincosf(0.5f * p0[off], &s_, &c_);
          ql = Quat{c_, 0.f, 0.f, s_};
          wl = v3(0.f, 0.f, w0[off]);
        }
        const Vec3 r = qrot(A[par].Q, v3(c.attach[j][0], c.attach[j][1], c.attach[j][2]));
        A[j].Q = qmul(A[par].Q, ql);
        A[j].p = A[par].p + r;
        A[j].w = A[par].w + qrot(A[j].Q, wl);
        A[j].v = A[par].v + cross(A[par].w, r);
      }
      const Vec3 rb = qrot(A[j].Q, v3(c.body_attach[j][0], c.body_attach[j][1], c.body_attach[j][2]));
      Vec3 pos = A[j].p + rb;                       // body centre (ground height 0)
      Vec3 lin = A[j].v + cross(A[j].w, rb);        // its linear velocity
      Vec3 ang = A[j].w;
      Vec3 nrm = qrot(A[j].Q, v3(0.f, 1.f, 0.f));   // CalcNormalTangent, MathUtil.cpp:622-628
      Vec3 tan = qrot(A[j].Q, v3(1.f, 0.f, 0.f));
      const bool is_root = j == 0;
      if (!f.all_world) {
        if (!f.world_root_pos || !is_root) {
          pos = rot_heading_inv(hc, hs, v3(pos.x - root.x, pos.y, pos.z - root.z));
          pos.y -= root.y;
        }
        if (!f.world_root_rot || !is_root) {
          nrm = rot_heading_inv(hc, hs, nrm);
          tan = rot_heading_inv(hc, hs, tan);
          lin = rot_heading_inv(hc, hs, lin);
          ang = rot_heading_inv(hc, hs, ang);
        }
      }
      float* po = o + 1 + 9 * j;
      po[0] = pos.x; po[1] = pos.y; po[2] = pos.z;
      po[3] = nrm.x; po[4] = nrm.y; po[5] = nrm.z;
      po[6] = tan.x; po[7] = tan.y; po[8] = tan.z;
      float* vo = o + vel0 + 6 * j;
      vo[0] = lin.x * f.vel_scale; vo[1] = lin.y * f.vel_scale; vo[2] = lin.z * f.vel_scale;
      vo[3] = ang.x * f.vel_scale; vo[4] = ang.y * f.vel_scale; vo[5] = ang.z * f.vel_scale;
    }
  }
}

}  // namespace simstep
