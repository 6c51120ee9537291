// Grouped, persistent, warp-specialised tcgen05 GEMM for the ensemble MLP layers
// and the random-feature projection:   D[g] = A[g] * B[g]^T  (both K-major).
//
//   * A (env rows x K) and B (out features x K) tiles arrive by TMA with the
//     128-byte swizzle, 4-stage mbarrier ring.
//   * One elected thread issues tcgen05.mma (M=128, N=256, fp32 accumulate in
//     TMEM); two 256-column accumulators are double-buffered so the epilogue
//     of tile i overlaps the MMAs of tile i+1.
//   * The K loop reads the first kb_x blocks from a tensor shared by all groups
//     (the normalised [s,a] input x) and the rest from the group's own
//     activation buffer: that IS the dense-connect concat of BasicMLP.forward
//     (reference milo/milo/dynamics.py:427-430) with no cat kernel.
//   * Epilogues: bias+activation written straight into the next layer's
//     K-slice (hidden), bias+un-normalise to fp32 (final, dynamics.py:231-232),
//     cos/dot for the RFF cost (linear_cost.py:64-71, 96-103).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace simstep {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kStages = 4;
constexpr int kNumEpiWarps = 4;
constexpr int kNumEpiThreads = kNumEpiWarps * 32;
constexpr int kGemmThreads = 64 + kNumEpiThreads;  // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int kTmemCols = 512;                     // two 128x256 fp32 accumulators
constexpr int kABytes = kBlockM * 128;             // one swizzle atom (128 B) per row
constexpr int kBBytes = kBlockN * 128;
constexpr int kStageBytes = kABytes + kBBytes;

// Operand formats.  cvt() is the scalar round-to-operand used by the pack kernels; store32<RELU>() converts
// 32 fp32 accumulator values (optionally through max(x, 0)) and writes them as one contiguous row chunk.
// NaN propagates through every variant (torch.relu(NaN) is NaN); fp16 saturates instead of overflowing.
struct ElemTF32 {
  using storage = float;
  static constexpr int kKind = 0;
  static constexpr uint32_t kFmt = 2;
  __device__ static __forceinline__ storage cvt(float x) { return ptx::round_tf32(x); }
  template <bool RELU>
  __device__ static __forceinline__ void store32(storage* dst, const float (&v)[32]) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float x = v[4 * j + i];
        if (RELU) asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(x) : "f"(x));
        r[i] = ptx::round_tf32(x);
      }
      d4[j] = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
};
struct ElemF16 {
  using storage = __half;
  static constexpr int kKind = 1;
  static constexpr uint32_t kFmt = 0;
  __device__ static __forceinline__ storage cvt(float x) {
    uint32_t r;  // saturate instead of overflowing to inf: an exploded state stays finite; NaN stays NaN
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(x));
    return __ushort_as_half(static_cast<unsigned short>(r & 0xFFFFu));
  }
  template <bool RELU>
  __device__ static __forceinline__ void store32(storage* dst, const float (&v)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // d = {a -> upper half, b -> lower half}
        if (RELU)
          asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(v[8 * j + 2 * i + 1]), "f"(v[8 * j + 2 * i]));
        else
          asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(v[8 * j + 2 * i + 1]), "f"(v[8 * j + 2 * i]));
      }
      d4[j] = make_uint4(p[0], p[1], p[2], p[3]);
    }
  }
};
struct ElemBF16 {
  using storage = __nv_bfloat16;
  static constexpr int kKind = 1;
  static constexpr uint32_t kFmt = 1;
  __device__ static __forceinline__ storage cvt(float x) { return __float2bfloat16_rn(x); }
  template <bool RELU>
  __device__ static __forceinline__ void store32(storage* dst, const float (&v)[32]) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (RELU)
          asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(v[8 * j + 2 * i + 1]), "f"(v[8 * j + 2 * i]));
        else
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p[i]) : "f"(v[8 * j + 2 * i + 1]), "f"(v[8 * j + 2 * i]));
      }
      d4[j] = make_uint4(p[0], p[1], p[2], p[3]);
    }
  }
};

template <typename E>
struct ElemDims {
  static constexpr int kBlockK = 128 / sizeof(typename E::storage);  // elements per swizzle row
  static constexpr int kUmmaK = 32 / sizeof(typename E::storage);    // K of one tcgen05.mma
};

enum EpiMode : int { kEpiHidden = 0, kEpiFinal = 1, kEpiRff = 2, kEpiHiddenTanh = 3 };

struct GemmArgs {
  // tile space: tile -> (group, m_tile, n_tile), n fastest
  int m_tiles;
  int n_tiles;
  int groups;
  // K loop
  int kb_x;               // k-blocks read through tmap_ax (rows shared by all groups)
  int kb_h0;              // first k-block inside tmap_ah
  int kb_h;               // k-blocks read through tmap_ah
  int a_rows_per_group;   // row stride between groups in tmap_ax / tmap_ah (x uses ax_rows_per_group)
  int ax_rows_per_group;  // 0 when x is shared
  int b_rows_per_group;
  // epilogue
  const float* bias;      // [groups][n_tiles*kBlockN] or nullptr
  const float* scale;     // kEpiFinal: [n_tiles*kBlockN] or nullptr; kEpiRff: w
  const float* shift;     // kEpiFinal: [n_tiles*kBlockN] or nullptr
  void* out;              // hidden: storage type; final: float; rff: phi float or nullptr
  long long out_pitch;    // elements
  long long out_group_stride;  // elements between groups
  int out_col0;
  int rows_valid;         // rows of the M axis (per group) that exist in `out`
  int cols_valid;
  int vec_ok;             // final: 16-byte aligned rows -> float4 stores
  float* rff_part;        // kEpiRff: [n_tiles][rff_part_stride] partial dots
  long long rff_part_stride;
  float rff_phi_scale;    // sqrt(2/D)
};

template <typename E>
__host__ __device__ constexpr uint32_t make_idesc() {
  return (1u << 4)                                   // D format: F32
         | (E::kFmt << 7) | (E::kFmt << 10)          // A / B format
         | (static_cast<uint32_t>(kBlockN >> 3) << 17)  // N
         | (static_cast<uint32_t>(kBlockM >> 4) << 24); // M
}

constexpr size_t gemm_smem_bytes() {
  return 1024 /*align slack*/ + size_t(kStages) * kStageBytes + 3 * kBlockN * sizeof(float) + 256;
}

// cos(x) for the random-feature epilogue: two-constant Cody-Waite reduction to [-pi, pi] then the SFU
// approximation (abs error < 1e-6 for |x| up to ~1e4, far inside the 1e-3 budget of the cost).
__device__ __forceinline__ float fast_cos(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(k, -6.28318548202514648f, x);   // 2*pi rounded to fp32
  r = fmaf(k, 1.74845553e-7f, r);                // 2*pi_fp32 - 2*pi
  return __cosf(r);
}

template <typename E, int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_ax, const __grid_constant__ CUtensorMap tmap_ah,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmArgs args) {
  using T = typename E::storage;
  constexpr int BK = ElemDims<E>::kBlockK;
  constexpr int UK = ElemDims<E>::kUmmaK;
  constexpr int kMmasPerBlock = BK / UK;  // 4

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_tiles = smem;                                           // kStages * (A | B)
  float* sm_bias = reinterpret_cast<float*>(smem + size_t(kStages) * kStageBytes);
  float* sm_scale = sm_bias + kBlockN;
  float* sm_shift = sm_scale + kBlockN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_shift + kBlockN);
  uint64_t* full_bar = bars;                   // [kStages]
  uint64_t* empty_bar = bars + kStages;        // [kStages]
  uint64_t* tmem_full_bar = bars + 2 * kStages;      // [2]
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2; // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kNumEpiThreads);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmap_ax);
    ptx::prefetch_tensormap(&tmap_ah);
    ptx::prefetch_tensormap(&tmap_b);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_base_smem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  const int total_tiles = args.m_tiles * args.n_tiles * args.groups;
  const int kb_total = args.kb_x + args.kb_h;

  if (warp == 0) {
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_tile = tile % args.n_tiles;
      const int t2 = tile / args.n_tiles;
      const int m_tile = t2 % args.m_tiles;
      const int g = t2 / args.m_tiles;
      const int row_ax = g * args.ax_rows_per_group + m_tile * kBlockM;
      const int row_ah = g * args.a_rows_per_group + m_tile * kBlockM;
      const int row_b = g * args.b_rows_per_group + n_tile * kBlockN;
      for (int kb = 0; kb < kb_total; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem_tiles + size_t(stage) * kStageBytes;
          uint8_t* sb = sa + kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
          if (kb < args.kb_x) {
            ptx::tma_load_2d(sa, &tmap_ax, &full_bar[stage], kb * BK, row_ax);
          } else {
            ptx::tma_load_2d(sa, &tmap_ah, &full_bar[stage], (args.kb_h0 + kb - args.kb_x) * BK, row_ah);
          }
          ptx::tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * BK, row_b);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc<E>();
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      ptx::tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kBlockN;
      for (int kb = 0; kb < kb_total; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t sa = ptx::smem_u32(smem_tiles + size_t(stage) * kStageBytes);
          const uint64_t da = ptx::umma_desc_k_sw128(sa);
          const uint64_t db = ptx::umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kMmasPerBlock; ++k) {
            // advance 32 bytes (= UMMA_K elements) along K inside the swizzle atom
            ptx::umma_ss<E::kKind>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (kb == kb_total - 1) ptx::umma_commit(&tmem_full_bar[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row_in_tile = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % args.n_tiles;
      const int t2 = tile / args.n_tiles;
      const int m_tile = t2 % args.m_tiles;
      const int g = t2 / args.m_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n0 = n_tile * kBlockN;

      // per-tile column vectors -> smem (previous tile's readers are past this barrier)
      ptx::named_bar_sync(1, kNumEpiThreads);
      for (int i = epi_tid; i < kBlockN; i += kNumEpiThreads) {
        sm_bias[i] = args.bias ? args.bias[size_t(g) * args.n_tiles * kBlockN + n0 + i] : 0.f;
        if constexpr (MODE != kEpiHidden) {
          sm_scale[i] = args.scale ? args.scale[n0 + i] : 1.f;
          sm_shift[i] = (MODE == kEpiFinal && args.shift) ? args.shift[n0 + i] : 0.f;
        }
      }
      ptx::named_bar_sync(1, kNumEpiThreads);

      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
      const long long row = static_cast<long long>(m_tile) * kBlockM + row_in_tile;
      constexpr int kChunks = kBlockN / 32;

      // Two register buffers: the TMEM load of chunk c+1 is in flight while chunk c is processed, and the
      // accumulator is handed back to the MMA warp as soon as the last chunk sits in registers.
      uint32_t ra[32], rb[32];
      ptx::tmem_ld_32x32(taddr, ra);

      float dot = 0.f;  // kEpiRff
      auto process = [&](const uint32_t (&r)[32], int c) {
        const float4* b4 = reinterpret_cast<const float4*>(sm_bias + c * 32);
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = b4[j];
          v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
          v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
          v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
          v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
        }
        if constexpr (MODE == kEpiHidden || MODE == kEpiHiddenTanh) {
          T* orow = static_cast<T*>(args.out) + size_t(g) * args.out_group_stride + row * args.out_pitch +
                    args.out_col0 + n0 + c * 32;
          if constexpr (MODE == kEpiHiddenTanh) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
            E::template store32<false>(orow, v);
          } else {
            E::template store32<true>(orow, v);
          }
        } else if constexpr (MODE == kEpiFinal) {
          float* orow = static_cast<float*>(args.out) + size_t(g) * args.out_group_stride + row * args.out_pitch +
                        args.out_col0 + n0 + c * 32;
          const float4* s4 = reinterpret_cast<const float4*>(sm_scale + c * 32);
          const float4* h4 = reinterpret_cast<const float4*>(sm_shift + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 sc = s4[j], sh = h4[j];
            v[4 * j + 0] = fmaf(v[4 * j + 0], sc.x, sh.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], sc.y, sh.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], sc.z, sh.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], sc.w, sh.w);
          }
          const int col = n0 + c * 32;
          if (row < args.rows_valid) {
            if (args.vec_ok && col + 32 <= args.cols_valid) {
              float4* dst = reinterpret_cast<float4*>(orow);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col + j < args.cols_valid) orow[j] = v[j];
            }
          }
        } else {  // kEpiRff: phi = cos(pre-activation); dot with w (padded columns carry w == 0)
          const float4* w4 = reinterpret_cast<const float4*>(sm_scale + c * 32);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fast_cos(v[j]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 w = w4[j];
            dot = fmaf(f[4 * j + 0], w.x, dot);
            dot = fmaf(f[4 * j + 1], w.y, dot);
            dot = fmaf(f[4 * j + 2], w.z, dot);
            dot = fmaf(f[4 * j + 3], w.w, dot);
          }
          if (args.out != nullptr && row < args.rows_valid) {
            float* prow = static_cast<float*>(args.out) + row * args.out_pitch + n0 + c * 32;
            const int col = n0 + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col + j < args.cols_valid) prow[j] = f[j] * args.rff_phi_scale;
          }
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
          ptx::mbar_arrive(&tmem_empty_bar[acc]);  // accumulator drained: the MMA warp may overwrite it
        }
        process(rb, c + 1);
      }
      if constexpr (MODE == kEpiRff) {
        if (args.rff_part != nullptr && row < args.rows_valid)
          args.rff_part[size_t(n_tile) * args.rff_part_stride + row] = dot;
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace simstep
