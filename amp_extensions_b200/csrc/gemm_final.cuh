// Final ensemble layer FUSED with the env step's elementwise tail: one launch computes, for every member, the
// un-normalised deltas  y_g = W5_g [x; h_1..h_4] + b5_g,  delta_g = y_g * sigma_d + mu_d  (reference
// milo/milo/dynamics.py:231-232, 427-433) on the tcgen05 tensor cores AND, tile by tile, the next state
// s' = s + delta_{m_e}, the step counter and termination mask (gym-simenv/gym_simenv/envs/sim_env.py:153-173,
// 175-268), the ensemble discrepancy  max_{i<j} ||delta_i - delta_j||_2  (dynamics.py:134-143) and the [s; s'] operand
// rows of the cost features (milo/milo/linear_cost.py:115-126).  It replaces the final-layer GEMM launch followed by
// post_step_tma_kernel: the member deltas no longer make an HBM round trip between two kernels.
//
// How the members of one env tile meet.  A 128-row x 226-column fp32 delta tile per member does not fit on chip
// next to the operand ring (4 x 113 KB), and 4 x 226 accumulator columns exceed the 512 TMEM columns, so the tile
// space stays (env tile, member) - 4x more work units than env tiles, which keeps the persistent grid's wave
// quantisation at 94 % - and the members meet in L2: every member tile's epilogue writes its deltas to a
// TRANSPOSED scratch block ([float4 column][row], coalesced 512-byte stores), then takes a ticket on the env
// block's counter.  The CTA that draws the LAST ticket runs the tail for that 128-row block: two threads per row, all N
// members' float4 columns read back with coalesced L2 loads (ld.global.cg), pair distances accumulated in registers,
// the state rows staged by ONE bulk copy per warp (16 rows are one contiguous, 16-byte granular span), s' formed in
// place and written back by one bulk store.  Nobody ever waits on another CTA (the last arriver does the work), so
// there is no co-residency assumption.  The scratch is written and re-read within about one tile time and stays in
// the 126 MB L2.
//
// The MMA pipeline never waits for the tail: the accumulator is handed back to the MMA warp before the ticket is
// taken, and a tail costs less than one tile's MMAs (2 304 x 256 x 256 MACs per CTA pair).
#pragma once
#include "elementwise.cuh"
#include "gemm_tcgen05.cuh"
#include "post_tma.cuh"

namespace simstep {

constexpr int kFinalStages = 5;     // operand ring: 5 x 32 KB (3 stages cannot cover the HBM latency: measured 2x slower)
constexpr int kFinalSlabRows = 16;  // one state slab per tail warp: 16 rows, two threads per row
constexpr int kFinalTailWarps = 4;  // warps 6..9: run the blocks' tails off the epilogue's critical path
constexpr int kFinalThreads = kGemmThreads + kFinalTailWarps * 32;
constexpr int kFinalQueue = 64;     // blocks a CTA may have queued for its tail warps (>= tiles per CTA pair)

struct FinalArgs {
  // tile space: tile -> (m_tile, group), group fastest, so the members of an env tile run side by side
  int m_tiles;            // 256-row tiles (CTA pairs)
  int groups;             // ensemble members
  int kb_x, kb_h0, kb_h;  // K loop, as GemmArgs
  int a_rows_per_group;
  int b_rows_per_group;
  const float* bias;      // [groups][kBlockN]
  const float* scale;     // [kBlockN] or nullptr (no output transform)
  const float* shift;
  // member deltas, transposed: [group][block][c4][128 rows] float4, block = 128 rows, c4 < ceil(S / 4)
  float4* dws_t;
  int n_blocks;            // 128-row blocks per group (row stride of the scratch)
  unsigned int* counters;  // [n_blocks] tickets, zero between launches
  // env step (any of these may be null: discrepancy-only / no counters / no mask)
  const float* state;      // [n_rows][S]
  float* next_state;       // [n_rows][S], may alias state
  const int32_t* member;
  int32_t* num_steps;
  float* disc;
  uint8_t* done;
  long long n_rows;
  int S;
  TermConst tc;
  // cost-feature operand rows [s | 0.. | s' | 0..] (+ the same again as low parts when split)
  void* rff_out;           // nullptr: not produced
  long long rff_pitch;     // elements per operand row
  int rff_col2;            // first column of s'
  int rff_lo_off;          // column offset of the low parts, 0: no split
  int debug;               // SIMSTEP_FINAL_DEBUG bit mask: timing experiments only (results are wrong when set)
};

// bias of every member + output scale / shift, staged once per CTA (the epilogue's broadcast reads would otherwise
// go through the ~4 KB of L1 this kernel's shared-memory footprint leaves)
constexpr int kFinalMaxGroups = 8;
inline size_t final_const_bytes(int groups) { return size_t(groups + 2) * kBlockN * sizeof(float); }
inline size_t final_smem_bytes(int S, int groups) {
  return 1024 + size_t(kFinalStages) * GemmShape<2>::kStageBytes + size_t(4) * kFinalSlabRows * S * sizeof(float) +
         final_const_bytes(groups) + 256 + (kFinalQueue + 4) * sizeof(int);
}

// 1-D bulk store shared -> global (16-byte aligned addresses, size a multiple of 16), bulk async group.
__device__ __forceinline__ void bulk_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :
               : "l"(reinterpret_cast<uint64_t>(gmem_dst)), "r"(ptx::smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

// 8 consecutive operand elements (16 bytes for the 2-byte formats, 32 for tf32); p is 16-byte aligned
template <typename E>
__device__ __forceinline__ void final_store8(typename E::storage* p, const float (&v)[8]) {
  using T = typename E::storage;
  if constexpr (sizeof(T) == 2) {
    using P = typename Pair<T>::type;
    union { P p2[4]; uint4 u; } w;
#pragma unroll
    for (int i = 0; i < 4; ++i) w.p2[i] = make_pair_cvt<E>(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = w.u;
  } else {
    float4 a, b;
    a.x = E::cvt(v[0]); a.y = E::cvt(v[1]); a.z = E::cvt(v[2]); a.w = E::cvt(v[3]);
    b.x = E::cvt(v[4]); b.y = E::cvt(v[5]); b.z = E::cvt(v[6]); b.w = E::cvt(v[7]);
    reinterpret_cast<float4*>(p)[0] = a;
    reinterpret_cast<float4*>(p)[1] = b;
  }
}

template <typename E>
__device__ __forceinline__ void final_rff_store(const FinalArgs& a, long long row, int col, const float (&v)[8]) {
  using T = typename E::storage;
  T* orow = static_cast<T*>(a.rff_out) + row * a.rff_pitch + col;
  final_store8<E>(orow, v);
  if (a.rff_lo_off) {
    float lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) lo[i] = v[i] - static_cast<float>(E::cvt(v[i]));
    final_store8<E>(orow + a.rff_lo_off, lo);
  }
}

// The tail of one 128-row block, run by the four epilogue warps of the CTA that took the block's last ticket, in two
// passes of 64 rows.  In a pass warp q owns 16 rows and its own slab; TWO threads share a row (lanes r and r + 16),
// each taking half of the row's float4 columns, so the per-thread dependent chain (L2 load -> distances -> s' ->
// operand stores) is half as long and a warp's scratch loads are two coalesced 256-byte segments.
// __noinline__: the tail gets its own register allocation; inlined, its live ranges made ptxas spill inside the
// per-tile epilogue loop (measured: the spills alone doubled the kernel's time - this kernel leaves ~4 KB of L1).
template <typename E, int NM>
__device__ __noinline__ void final_tail(const FinalArgs& a, int block, int pass, int q, int lane, float* slab,
                                        uint64_t* slab_bar, uint32_t& slab_phase, bool& store_pending) {
  constexpr int NP = NM * (NM - 1) / 2;
  const int S = a.S;
  const int n_c4 = (S + 3) >> 2;
  const int r = lane & (kFinalSlabRows - 1);
  const int half = lane >> 4;
  const int row_in_block = pass * (4 * kFinalSlabRows) + q * kFinalSlabRows;  // first row of this warp's slab
  const long long row0 = static_cast<long long>(block) * kBlockM + row_in_block;
  const long long row = row0 + r;
  const bool valid = row < a.n_rows;
  const long long left = a.n_rows - row0;
  const int rows_here = left <= 0 ? 0 : (left >= kFinalSlabRows ? kFinalSlabRows : int(left));
  if (rows_here == 0) return;  // warp-uniform: nothing of this slab exists
  const int even_rows = rows_here & ~1;
  const bool have_next = a.next_state != nullptr && !(a.debug & 16);

  if (have_next) {
    if (store_pending) {  // the previous bulk store must have finished reading the slab
      if (lane == 0) ptx::tma_store_wait_read<0>();
      __syncwarp();
      store_pending = false;
    }
    if (lane == 0) {
      if (even_rows) {
        const uint32_t bytes = uint32_t(even_rows) * S * 4;
        ptx::mbar_arrive_expect_tx(slab_bar, bytes);
        ptx::bulk_load_1d(slab, a.state + row0 * S, bytes, slab_bar);
      } else {
        ptx::mbar_arrive(slab_bar);
      }
    }
    if (rows_here & 1) {  // an odd trailing row is not a bulk-copy size: plain loads
      const float* src = a.state + (row0 + even_rows) * S;
      for (int k = lane; k < S; k += 32) slab[even_rows * S + k] = src[k];
    }
  }
  int mem = 0, steps = 0;
  if (valid && a.member != nullptr) mem = __ldg(a.member + row);
  if (valid && half == 0 && a.num_steps != nullptr) steps = a.num_steps[row] + 1;

  // this thread's share of the row's float4 columns: [c4_lo, c4_hi), c4_lo even (16-byte operand stores)
  const int c4_split = ((n_c4 >> 1) + 1) & ~1;
  const int c4_lo = half ? c4_split : 0;
  const int c4_hi = half ? n_c4 : (c4_split < n_c4 ? c4_split : n_c4);
  const size_t gstride = size_t(a.n_blocks) * n_c4 * kBlockM;
  // plain (weak) loads: the acquire fence after the ticket invalidated L1, and nothing writes the scratch any more,
  // so the compiler is free to batch the loads of several iterations (the tail is bound by L2 latency, not issue)
  const float4* __restrict__ dp = a.dws_t + size_t(block) * n_c4 * kBlockM + (row_in_block + r);
  float* srow = slab + r * S;
  float acc[NP > 0 ? NP : 1];
#pragma unroll
  for (int p = 0; p < (NP > 0 ? NP : 1); ++p) acc[p] = 0.f;
  bool vel_flag = false;
  constexpr PostPairs<NM> pairs{};
  const bool want_rff = have_next && a.rff_out != nullptr && valid && !(a.debug & 4);

  if (have_next) {
    ptx::mbar_wait(slab_bar, slab_phase);
    slab_phase ^= 1;
    __syncwarp();
  }

  // 8 consecutive columns of this thread's slab row (zeros beyond S; S is even, so pairs are valid as a whole)
  auto slab_load8 = [&](int col, float (&v)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      if (col + k < S) {
        const float2 t = *reinterpret_cast<const float2*>(srow + col + k);
        v[k] = t.x;
        v[k + 1] = t.y;
      } else {
        v[k] = v[k + 1] = 0.f;
      }
    }
  };

  // phase 0: the s half of the cost-feature operand row (before s' overwrites the slab)
  if (want_rff) {
#pragma unroll 2
    for (int c4 = c4_lo; c4 < c4_hi; c4 += 2) {
      float sv[8];
      slab_load8(4 * c4, sv);
      final_rff_store<E>(a, row, 4 * c4, sv);
    }
  }

  // phase 1: member deltas from the scratch -> pair distances, s' in place.  No global stores in this loop, and the
  // loads run up to three iterations (24 x 16 bytes per thread) ahead of their use through rotating register sets:
  // with one set the loop paid a full L2 round trip per iteration (measured: the tail took ~3x a tile's MMAs).
  auto load_cols = [&](float4 (&u)[NM][2], int c4) {
    const int c4b = c4 + 1 < c4_hi ? c4 + 1 : c4;
#pragma unroll
    for (int g = 0; g < NM; ++g) {
      u[g][0] = dp[g * gstride + size_t(c4) * kBlockM];
      u[g][1] = dp[g * gstride + size_t(c4b) * kBlockM];
    }
  };
  auto consume = [&](const float4 (&u)[NM][2], int c4) {
    const bool has1 = c4 + 1 < c4_hi;
    float d[NM][8];
#pragma unroll
    for (int g = 0; g < NM; ++g) {
      d[g][0] = u[g][0].x; d[g][1] = u[g][0].y; d[g][2] = u[g][0].z; d[g][3] = u[g][0].w;
      d[g][4] = has1 ? u[g][1].x : 0.f; d[g][5] = has1 ? u[g][1].y : 0.f;
      d[g][6] = has1 ? u[g][1].z : 0.f; d[g][7] = has1 ? u[g][1].w : 0.f;
    }
    // padded output columns (>= S) carry 0 in every member (zero weights, zero bias, unit scale), so they add
    // nothing to the distances
    if (a.disc != nullptr) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int i0 = pairs.a[p], i1 = pairs.b[p];
        float s2 = acc[p];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float t = d[i0][k] - d[i1][k];
          s2 = fmaf(t, t, s2);
        }
        acc[p] = s2;
      }
    }
    if (have_next) {
      const int col = 4 * c4;
      float sv[8], nx[8];
      slab_load8(col, sv);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // a member index outside [0, N) (the reference raises IndexError) makes the row's next state NaN
        float dm = __int_as_float(0x7fc00000);
#pragma unroll
        for (int g = 0; g < NM; ++g) dm = (g == mem) ? d[g][k] : dm;
        nx[k] = sv[k] + dm;
      }
#pragma unroll
      for (int k = 0; k < 8; k += 2)
        if (col + k < S) *reinterpret_cast<float2*>(srow + col + k) = make_float2(nx[k], nx[k + 1]);
      if (a.tc.enable_velocity_check) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          vel_flag = vel_flag || (col + k < S && PostVec<1>::vel_over(nx[k], col + k, a.tc.vel_offset,
                                                                      a.tc.vel_inv_divisor, a.tc.vel_threshold));
      }
    }
  };
  if (!(a.debug & 8)) {
    constexpr int kDepth = NM <= 4 ? 3 : 2;  // register sets in rotation (NM x 8 registers each)
    float4 u[kDepth][NM][2];
#pragma unroll
    for (int i = 0; i < kDepth - 1; ++i)
      if (c4_lo + 2 * i < c4_hi) load_cols(u[i], c4_lo + 2 * i);
#pragma unroll 1
    for (int c4 = c4_lo; c4 < c4_hi; c4 += 2 * kDepth) {
#pragma unroll
      for (int i = 0; i < kDepth; ++i) {
        const int cc = c4 + 2 * i;
        if (cc < c4_hi) {
          const int pre = cc + 2 * (kDepth - 1);
          if (pre < c4_hi) load_cols(u[(i + kDepth - 1) % kDepth], pre);
          consume(u[i], cc);
        }
      }
    }
  }

  // phase 2: the s' half of the operand row
  if (want_rff) {
#pragma unroll 2
    for (int c4 = c4_lo; c4 < c4_hi; c4 += 2) {
      float nx[8];
      slab_load8(4 * c4, nx);
      final_rff_store<E>(a, row, a.rff_col2 + 4 * c4, nx);
    }
  }

  if (a.disc != nullptr) {
    float best = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const float s2 = acc[p] + __shfl_xor_sync(0xffffffffu, acc[p], 16);  // the two halves of the row
      if (s2 != s2) best = s2;  // NaN propagates like torch.max
      else if (best == best && s2 > best) best = s2;
    }
    if (valid && half == 0) a.disc[row] = sqrtf(best);
  }
  if (have_next) {
    vel_flag = __shfl_xor_sync(0xffffffffu, vel_flag ? 1 : 0, 16) != 0 || vel_flag;
    __syncwarp();  // both halves of every row are in the slab
    if (valid && half == 0) {
      if (a.num_steps != nullptr) a.num_steps[row] = steps;
      if (a.done != nullptr) {
        bool flag = vel_flag;
        const TermConst& tc = a.tc;
        for (int b = 0; b < tc.n_bodies; ++b) {
          const int off = tc.body_offset[b];
          const int shape = tc.body_shape[b];
          float y = srow[off + 1];
          if (!(tc.record_all_world || (b == 0 && tc.record_world_root_pos))) y += srow[0];
          const float lim = tc.body_radius[b] + 0.0001f;
          if (shape == SIMSTEP_SHAPE_SPHERE) {
            flag = flag || (y <= lim);
          } else if (shape == SIMSTEP_SHAPE_CAPSULE) {
            const float cap = tc.body_half_height[b] * srow[off + tc.pos_dim + 1];
            flag = flag || (y + cap <= lim) || (y - cap <= lim);
          }
        }
        a.done[row] = (flag || (a.num_steps != nullptr && steps >= tc.horizon)) ? 1 : 0;
      }
    }
    // s' rows leave through the async proxy: order this warp's shared-memory writes before it
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && even_rows) {
      bulk_store_1d(a.next_state + row0 * S, slab, uint32_t(even_rows) * S * 4);
      ptx::tma_store_commit();
    }
    store_pending = true;
    if (rows_here & 1) {
      float* dst = a.next_state + (row0 + even_rows) * S;
      for (int k = lane; k < S; k += 32) dst[k] = slab[even_rows * S + k];
    }
  }
}

template <typename E, int NM>
__global__ void __launch_bounds__(kFinalThreads, 1)
final_fused_kernel(const __grid_constant__ CUtensorMap tmap_ax, const __grid_constant__ CUtensorMap tmap_ah,
                   const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ FinalArgs args) {
  constexpr int CG = 2;
  using Sh = GemmShape<CG>;
  constexpr int BK = ElemDims<E>::kBlockK;
  constexpr int UK = ElemDims<E>::kUmmaK;
  constexpr int kMmasPerBlock = BK / UK;
  constexpr int kStages = kFinalStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_tiles = smem;
  float* smem_slabs = reinterpret_cast<float*>(smem + size_t(kStages) * Sh::kStageBytes);
  float* smem_const = smem_slabs + size_t(4) * kFinalSlabRows * args.S;  // [groups] bias | scale | shift
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_const + size_t(args.groups + 2) * kBlockN);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint64_t* slab_bar = bars + 2 * kStages + 4;  // [4]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 8);
  // hand-over of completed blocks from the epilogue (thread 64 draws the tickets) to the tail warps
  volatile int* q_head = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 256);
  volatile int* q_finished = q_head + 1;
  volatile int* q_blocks = q_head + 4;  // [kFinalQueue]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full_bar[i], 1);
      ptx::mbar_init(&tmem_empty_bar[i], kNumEpiWarps * CG);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&slab_bar[i], 1);
    *q_head = 0;
    *q_finished = 0;
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmap_ax);
    ptx::prefetch_tensormap(&tmap_ah);
    ptx::prefetch_tensormap(&tmap_b);
  }
  if (warp == 1) {
    ptx::tmem_alloc<CG>(tmem_base_smem, kTmemCols);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  ptx::grid_dep_wait();
  ptx::grid_dep_launch();

  const int total_tiles = args.m_tiles * args.groups;
  const int kb_total = args.kb_x + args.kb_h;
  const int tile0 = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;

  if (warp == 0) {
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int g = tile % args.groups;
      const int m_tile = tile / args.groups;
      const int m_row = (m_tile * CG + int(cta_rank)) * kBlockM;
      const int row_ah = g * args.a_rows_per_group + m_row;
      const int row_b = g * args.b_rows_per_group + int(cta_rank) * Sh::kBRows;
      for (int kb = 0; kb < kb_total; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem_tiles + size_t(stage) * Sh::kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], Sh::kStageBytes * CG);
          if (kb < args.kb_x) {
            ptx::tma_load_2d<CG>(sa, &tmap_ax, &full_bar[stage], kb * BK, m_row);
          } else {
            ptx::tma_load_2d<CG>(sa, &tmap_ah, &full_bar[stage], (args.kb_h0 + kb - args.kb_x) * BK, row_ah);
          }
          ptx::tma_load_2d<CG>(sb, &tmap_b, &full_bar[stage], kb * BK, row_b);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc<E, CG>();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          if (lane == 0) {
            const uint32_t sa = ptx::smem_u32(smem_tiles + size_t(stage) * Sh::kStageBytes);
            const uint64_t da = ptx::umma_desc_k_sw128(sa);
            const uint64_t db = ptx::umma_desc_k_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kMmasPerBlock; ++k)
              ptx::umma_ss<E::kKind, CG>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            ptx::umma_commit<CG>(&empty_bar[stage]);
            if (kb == kb_total - 1) ptx::umma_commit<CG>(&tmem_full_bar[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue warps: deltas -> transposed scratch, ticket =====
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    const int n_c4 = (args.S + 3) >> 2;
    int q_pushed = 0;
    for (int i = epi_tid; i < (args.groups + 2) * kBlockN; i += kNumEpiThreads) {
      const int gi = i / kBlockN, ci = i % kBlockN;
      float v;
      if (gi < args.groups) v = args.bias[size_t(gi) * kBlockN + ci];
      else if (gi == args.groups) v = args.scale ? args.scale[ci] : 1.f;
      else v = args.shift ? args.shift[ci] : 0.f;
      smem_const[i] = v;
    }
    ptx::named_bar_sync(1, kNumEpiThreads);
    const float* sm_scale = smem_const + size_t(args.groups) * kBlockN;
    const float* sm_shift = sm_scale + kBlockN;
    int it = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
      const int g = tile % args.groups;
      const int m_tile = tile / args.groups;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int block = m_tile * CG + int(cta_rank);
      const float* bias_g = smem_const + size_t(g) * kBlockN;
      float4* dst = args.dws_t + (size_t(g) * args.n_blocks + block) * n_c4 * kBlockM + row_in_tile;

      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
      constexpr int kChunks = kBlockN / 32;

      uint32_t ra[32], rb[32];
      ptx::tmem_ld_32x32(taddr, ra);
      auto process = [&](const uint32_t (&r)[32], int c) {
        if (c * 8 >= n_c4) return;  // chunk holds padded columns only
        float v[32];
        const float4* b4 = reinterpret_cast<const float4*>(bias_g + c * 32);
        const float4* s4 = reinterpret_cast<const float4*>(sm_scale + c * 32);
        const float4* h4 = reinterpret_cast<const float4*>(sm_shift + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = b4[j], sc = s4[j], sh = h4[j];  // same address in every lane: shared-memory broadcast
          v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]) + b.x, sc.x, sh.x);
          v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]) + b.y, sc.y, sh.y);
          v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]) + b.z, sc.z, sh.z);
          v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]) + b.w, sc.w, sh.w);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c4 = c * 8 + j;
          if (c4 < n_c4 && !(args.debug & 2)) __stcg(dst + size_t(c4) * kBlockM, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
      };
#pragma unroll 1
      for (int c = 0; c < kChunks; c += 2) {
        ptx::tmem_ld_wait();
        ptx::tmem_ld_32x32(taddr + (c + 1) * 32, rb);
        process(ra, c);
        ptx::tmem_ld_wait();
        if (c + 2 < kChunks) {
          ptx::tmem_ld_32x32(taddr + (c + 2) * 32, ra);
        } else {
          ptx::tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster<CG>(&tmem_empty_bar[acc], 0);
        }
        process(rb, c + 1);
      }

      // ticket: the member tile whose arrival completes the block queues the block's tail for this CTA's tail warps
      if (args.debug & 1) continue;
      asm volatile("fence.acq_rel.gpu;" ::: "memory");  // release this thread's scratch stores
      ptx::named_bar_sync(1, kNumEpiThreads);
      if (epi_tid == 0) {
        const unsigned int old = atomicAdd(args.counters + block, 1u);
        if (old == static_cast<unsigned int>(args.groups - 1)) {
          args.counters[block] = 0;  // nobody else touches it again in this launch
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          q_blocks[q_pushed % kFinalQueue] = block;
          __threadfence_block();
          *q_head = ++q_pushed;
        }
      }
    }
    if (epi_tid == 0) {
      __threadfence_block();
      *q_finished = 1;
    }
  } else if (warp < 6 + kFinalTailWarps) {
    // ===== tail warps: the env step's tail of every block this CTA completed =====
    const int tq = warp - 6;
    float* slab = smem_slabs + size_t(tq) * kFinalSlabRows * args.S;
    uint32_t slab_phase = 0;
    bool store_pending = false;
    int next = 0;
    long long t_idle = clock64();
    while (true) {
      int head = *q_head;
      if (head <= next) {
        if (*q_finished) {
          head = *q_head;  // the flag is raised after the last push
          if (head <= next) break;
        } else {
          __nanosleep(200);
          if (clock64() - t_idle > 8000000000LL) {  // a protocol bug must not hang the GPU: fail the launch instead
            if (lane == 0) printf("simstep: tail warp %d of block %d starved\n", tq, blockIdx.x);
            __trap();
          }
          continue;
        }
      }
      t_idle = clock64();
      const int block = q_blocks[next % kFinalQueue];
      ++next;
      // acquire: the other CTAs' scratch stores are ordered before their tickets, the last ticket before the
      // queue entry; at gpu scope the fence also drops this SM's stale L1 lines
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      if (args.debug & 64) continue;
      final_tail<E, NM>(args, block, 0, tq, lane, slab, &slab_bar[tq], slab_phase, store_pending);
      final_tail<E, NM>(args, block, 1, tq, lane, slab, &slab_bar[tq], slab_phase, store_pending);
    }
    if (store_pending && lane == 0) ptx::tma_store_wait<0>();
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

}  // namespace simstep
