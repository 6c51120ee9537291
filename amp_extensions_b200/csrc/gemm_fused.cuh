// All layers of the ensemble forward pass in ONE persistent launch (reference milo/milo/dynamics.py:422-433,
// BasicMLP.forward for every member): the per-layer kernel of gemm_tcgen05.cuh with its tile space extended over
// the layers,
//     tile -> (layer, member, m_tile, n_tile),   layers in order, n fastest,
// and the dense-connect data dependency carried at tile granularity instead of at kernel boundaries: a tile of
// layer l reads the rows of its (member, m_tile) that EVERY n-tile of layer l-1 wrote, so each CTA bumps
// ready[l-1][member][m_tile] (release, after its TMA stores have completed) and the TMA producer of a layer-l tile
// spins on that counter (acquire) before it loads the first k-block that comes from the activation buffer.  CTA
// pairs take tiles round-robin in increasing order, a tile only waits on tiles with smaller indices, and all 148
// CTAs are co-resident (one per SM): the wait graph is well-founded, so there is no deadlock.
//
// What it is meant to buy: no drain / launch / pipeline-fill between layers, and at small batches layers that
// overlap instead of five latency-bound launches.  STATUS (round 1): opt-in (SIMSTEP_FUSED_LAYERS=1), results
// bit-identical to the per-layer path, but at the 40 000-env bench batch it needs 0.75 ms where the five per-layer
// launches need 0.68 ms (tensor pipe 62 % vs 75 % active under ncu): with the dependency tracking switched off the
// tile loop alone takes 0.70 ms (programmatic dependent launch already hides most of what a kernel boundary
// costs), the counters add 0.06 ms.  It is the faster path only for small eager batches.  The per-layer path stays the default until the gap is understood.
// Pipeline, barriers, TMEM double-buffering and the TMA-store epilogue are those of gemm_tcgen05.cuh (CG = 2 only).
#pragma once
#include "gemm_tcgen05.cuh"

namespace simstep {

constexpr int kFusedMaxLayers = SIMSTEP_MAX_HIDDEN + 1;

struct FusedLayer {
  int tile0;       // first tile index of the layer
  int n_tiles;     // n-tiles per (member, m_tile)
  int kb_x, kb_h0, kb_h;
  int b_rows_per_group;
  int out_col0;
  int mode;        // kEpiHidden, kEpiHiddenTanh or kEpiFinal
  const float* bias;
};

struct FusedArgs {
  int n_layers;
  int total_tiles;
  int m_tiles, groups;
  int a_rows_per_group;    // row stride between members in the activation buffer
  int ax_rows_per_group;   // 0: x shared by all members
  int out_rows_per_group;
  const float* scale;      // final layer: [SP] or nullptr
  const float* shift;
  unsigned int* ready;     // [n_layers][groups][m_tiles], zeroed before the launch
  FusedLayer layer[kFusedMaxLayers];
};

struct FusedMaps {
  CUtensorMap x, h, out_final;
  CUtensorMap w[kFusedMaxLayers];
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() {
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

struct FusedTile {
  int layer, g, m_tile, n_tile;
};
__device__ __forceinline__ FusedTile fused_decode(const FusedArgs& a, int tile) {
  int l = 0;
#pragma unroll 1
  while (l + 1 < a.n_layers && tile >= a.layer[l + 1].tile0) ++l;
  const int t = tile - a.layer[l].tile0;
  FusedTile r;
  r.layer = l;
  r.n_tile = t % a.layer[l].n_tiles;
  const int t2 = t / a.layer[l].n_tiles;
  r.m_tile = t2 % a.m_tiles;
  r.g = t2 / a.m_tiles;
  return r;
}

template <typename E>
__global__ void __launch_bounds__(kGemmThreads, 1)
ensemble_fused_kernel(const __grid_constant__ FusedMaps maps, const __grid_constant__ FusedArgs args) {
  constexpr int CG = 2;
  using S = GemmShape<CG>;
  constexpr int BK = ElemDims<E>::kBlockK;
  constexpr int UK = ElemDims<E>::kUmmaK;
  constexpr int kMmasPerBlock = BK / UK;
  constexpr int kStages = S::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_tiles = smem;
  uint8_t* smem_out = smem + size_t(kStages) * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + size_t(kOutStages) * kOutStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full_bar = bars + 2 * kStages;
  uint64_t* tmem_empty_bar = bars + 2 * kStages + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool leader = cta_rank == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kNumEpiWarps * CG);
    }
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&maps.x);
    ptx::prefetch_tensormap(&maps.h);
    ptx::prefetch_tensormap(&maps.out_final);
    for (int l = 0; l < args.n_layers; ++l) ptx::prefetch_tensormap(&maps.w[l]);
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

  const int tile_first = blockIdx.x / CG;
  const int tile_step = gridDim.x / CG;

  if (warp == 0) {
    // ===== TMA producer =====
    // The dependency counter of a tile is read one tile AHEAD (a relaxed load whose value is only looked at when
    // the tile starts, ~a tile's worth of MMAs later), so that in the common case - the counter was already
    // complete - the producer pays an acquire fence instead of an L2 round trip and keeps its stage look-ahead.
    auto flag_of = [&](const FusedTile& t) {
      return args.ready + (static_cast<size_t>(t.layer - 1) * args.groups + t.g) * args.m_tiles + t.m_tile;
    };
    auto needs_flag = [&](const FusedTile& t) { return t.layer > 0 && args.layer[t.layer].kb_h > 0; };
    int stage = 0;
    uint32_t phase = 0;
    unsigned int seen = 0;   // counter value prefetched for the CURRENT tile (lane 1)
    if (tile_first < args.total_tiles) {
      const FusedTile t0 = fused_decode(args, tile_first);
      if (lane == 1 && needs_flag(t0)) seen = ld_relaxed_gpu(flag_of(t0));
    }
    for (int tile = tile_first; tile < args.total_tiles; tile += tile_step) {
      const FusedTile t = fused_decode(args, tile);
      unsigned int seen_next = 0;
      if (tile + tile_step < args.total_tiles) {
        const FusedTile tn = fused_decode(args, tile + tile_step);
        if (lane == 1 && needs_flag(tn)) seen_next = ld_relaxed_gpu(flag_of(tn));
      }
      const FusedLayer& ly = args.layer[t.layer];
      const int m_row = (t.m_tile * CG + int(cta_rank)) * kBlockM;
      const int row_ax = t.g * args.ax_rows_per_group + m_row;
      const int row_ah = t.g * args.a_rows_per_group + m_row;
      const int row_b = t.g * ly.b_rows_per_group + t.n_tile * kBlockN + int(cta_rank) * S::kBRows;
      const int kb_total = ly.kb_x + ly.kb_h;
      bool deps_ok = !needs_flag(t);
      for (int kb = 0; kb < kb_total; ++kb) {
        if (kb >= ly.kb_x && !deps_ok) {
          // the previous layer's rows of this (member, m_tile): every n-tile of both CTAs of every pair that owned
          // one must have completed its stores
          const unsigned int want = static_cast<unsigned int>(args.layer[t.layer - 1].n_tiles) * CG;
          // Lane 1 does the acquire, lane 0 issues the copies: a gpu-scope fence waits for the executing thread's own
          // outstanding memory operations, and lane 0 has up to kStages TMA loads in flight that must not be drained
          // at every tile.  __syncwarp orders lane 1's acquire before lane 0's later loads.
          if (lane == 1) {
            if (seen < want) {
              // bounded spin: a dependency that never arrives is a bug, and a trap beats a hung GPU
              const unsigned int* flag = flag_of(t);
              unsigned int spins = 0;
              while (ld_acquire_gpu(flag) < want) {
                __nanosleep(64);
                if (++spins > (1u << 25)) __trap();
              }
            } else {
              fence_acq_rel_gpu();          // relaxed load + acquire fence = acquire
            }
          }
          __syncwarp();
          if (lane == 0) fence_proxy_async_global();   // the acquired data is read by the async proxy (TMA) next
          deps_ok = true;
        }
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem_tiles + size_t(stage) * S::kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes * CG);
          if (kb < ly.kb_x) {
            ptx::tma_load_2d<CG>(sa, &maps.x, &full_bar[stage], kb * BK, row_ax);
          } else {
            ptx::tma_load_2d<CG>(sa, &maps.h, &full_bar[stage], (ly.kb_h0 + kb - ly.kb_x) * BK, row_ah);
          }
          ptx::tma_load_2d<CG>(sb, &maps.w[t.layer], &full_bar[stage], kb * BK, row_b);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      seen = seen_next;
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {
      constexpr uint32_t idesc = make_idesc<E, CG>();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile_first; tile < args.total_tiles; tile += tile_step, ++it) {
        const FusedTile t = fused_decode(args, tile);
        const int kb_total = args.layer[t.layer].kb_x + args.layer[t.layer].kb_h;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kBlockN;
        for (int kb = 0; kb < kb_total; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tcgen05_fence_after();
          if (lane == 0) {
            const uint32_t sa = ptx::smem_u32(smem_tiles + size_t(stage) * S::kStageBytes);
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
  } else {
    // ===== epilogue warps =====
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const int epi_tid = threadIdx.x - 64;
    using T = typename E::storage;
    int store_it = 0;
    int it = 0;
    // Publishing a tile needs its stores COMPLETE.  When this CTA's next tile belongs to the same layer (it cannot
    // depend on this one) the wait is deferred until the next tile's first store group has been committed: by then
    // this tile's stores finished long ago and wait_group<1> returns at once.
    // The release itself is done by the neighbouring lane (epi_tid 1): a gpu-scope release waits for the executing
    // thread's own outstanding memory operations, and epi_tid 0 always has TMA stores in flight.
    unsigned int* pending = nullptr;   // warp-uniform in epilogue warp 0 (epi_tid < 32), nullptr elsewhere
    auto publish = [&](unsigned int* flag, bool all_groups) {   // called by the whole first epilogue warp
      if (epi_tid == 0) {
        if (all_groups) ptx::tma_store_wait<0>(); else ptx::tma_store_wait<1>();
        fence_proxy_async_global();
      }
      __syncwarp();
      if (epi_tid == 1) red_release_gpu_add(flag, 1u);
    };
    auto publish_pending_after_commit = [&]() {
      if (epi_tid < 32 && pending != nullptr) {
        publish(pending, false);
        pending = nullptr;
      }
    };
    for (int tile = tile_first; tile < args.total_tiles; tile += tile_step, ++it) {
      const FusedTile t = fused_decode(args, tile);
      const FusedLayer& ly = args.layer[t.layer];
      const bool is_final = ly.mode == kEpiFinal;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n0 = t.n_tile * kBlockN;
      const int m_row = (t.m_tile * CG + int(cta_rank)) * kBlockM;
      const float* bias_g = ly.bias ? ly.bias + size_t(t.g) * ly.n_tiles * kBlockN + n0 : nullptr;
      const CUtensorMap* tmap_out = is_final ? &maps.out_final : &maps.h;
      const int out_row = t.g * args.out_rows_per_group + m_row;

      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * kBlockN;
      constexpr int kChunks = kBlockN / 32;
      uint32_t ra[32], rb[32];
      ptx::tmem_ld_32x32(taddr, ra);

      // one 32-column chunk: bias (+ scale/shift | activation + convert), staged for a 128-byte-row TMA store
      auto process = [&](const uint32_t (&r)[32], int c) {
        float v[32];
        if (bias_g != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_g + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (is_final) {
          constexpr int kTileCols = 32;
          if (args.scale != nullptr) {
            const float4* s4 = reinterpret_cast<const float4*>(args.scale + n0 + c * 32);
            const float4* h4 = reinterpret_cast<const float4*>(args.shift + n0 + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 sc = __ldg(s4 + j), sh = __ldg(h4 + j);
              v[4 * j + 0] = fmaf(v[4 * j + 0], sc.x, sh.x);
              v[4 * j + 1] = fmaf(v[4 * j + 1], sc.y, sh.y);
              v[4 * j + 2] = fmaf(v[4 * j + 2], sc.z, sh.z);
              v[4 * j + 3] = fmaf(v[4 * j + 3], sc.w, sh.w);
            }
          }
          const int buf = store_it % kOutStages;
          if (epi_tid == 0) ptx::tma_store_wait_read<kOutStages - 1>();
          ptx::named_bar_sync(1, kNumEpiThreads);
          uint8_t* srow = smem_out + size_t(buf) * kOutStageBytes + row_in_tile * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(srow + ((u ^ (row_in_tile & 7)) << 4)) =
                make_uint4(__float_as_uint(v[4 * u]), __float_as_uint(v[4 * u + 1]), __float_as_uint(v[4 * u + 2]),
                           __float_as_uint(v[4 * u + 3]));
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(2, kNumEpiThreads);
          if (epi_tid == 0) {
            ptx::tma_store_2d(tmap_out, smem_out + size_t(buf) * kOutStageBytes, ly.out_col0 + n0 + c * kTileCols, out_row);
            ptx::tma_store_commit();
          }
          publish_pending_after_commit();
          ++store_it;
        } else {
          constexpr int kTileCols = int(128 / sizeof(T));
          constexpr int kChunksPerStore = kTileCols / 32;
          constexpr int kWords = E::kWords32;
          uint32_t w[kWords];
          if (ly.mode == kEpiHiddenTanh) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
            E::template pack32<false>(w, v);
          } else {
            E::template pack32<true>(w, v);
          }
          const int sub = c % kChunksPerStore;
          const int buf = store_it % kOutStages;
          if (sub == 0) {
            if (epi_tid == 0) ptx::tma_store_wait_read<kOutStages - 1>();
            ptx::named_bar_sync(1, kNumEpiThreads);
          }
          uint8_t* srow = smem_out + size_t(buf) * kOutStageBytes + row_in_tile * 128;
          constexpr int kUnits = kWords / 4;
#pragma unroll
          for (int u = 0; u < kUnits; ++u) {
            const int unit = sub * kUnits + u;
            *reinterpret_cast<uint4*>(srow + ((unit ^ (row_in_tile & 7)) << 4)) =
                make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
          }
          if (sub == kChunksPerStore - 1) {
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(2, kNumEpiThreads);
            if (epi_tid == 0) {
              ptx::tma_store_2d(tmap_out, smem_out + size_t(buf) * kOutStageBytes,
                                ly.out_col0 + n0 + (c / kChunksPerStore) * kTileCols, out_row);
              ptx::tma_store_commit();
            }
            publish_pending_after_commit();
            ++store_it;
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
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster<CG>(&tmem_empty_bar[acc], 0);
        }
        process(rb, c + 1);
      }
      // this CTA's half of the tile is on its way: publish it to the producers of the next layer once the stores have
      // COMPLETED (not only been read out of shared memory) - now if this CTA's next tile may depend on it
      if (!is_final && epi_tid < 32) {
        unsigned int* flag = args.ready + (static_cast<size_t>(t.layer) * args.groups + t.g) * args.m_tiles + t.m_tile;
        const int next = tile + tile_step;
        const bool next_same_layer = next < args.total_tiles && (t.layer + 1 >= args.n_layers || next < args.layer[t.layer + 1].tile0);
        if (next_same_layer) pending = flag;
        else publish(flag, true);
      }
    }
    if (epi_tid == 0) ptx::tma_store_wait<0>();
  }

  ptx::tcgen05_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

}  // namespace simstep
