// Persistent warp-specialised tcgen05 GEMM skeleton (one CTA per SM, 192 threads):
//   warp 0  lane 0 : TMA producer      (full/empty mbarrier ring over STAGES smem stages)
//   warp 1  lane 0 : MMA issuer        (tcgen05.mma, accumulators double-buffered in TMEM)
//   warps 2..5     : epilogue          (tcgen05.ld -> registers -> policy epilogue -> global)
// A tile is always 128 rows (UMMA M = 128, one TMEM lane per row) x 64 fp16 of K per stage
// (one SWIZZLE_128B row per tile row).  The policy P supplies the problem-specific parts:
//   P::BN, P::STAGES, P::B_MN_MAJOR, P::Params, P::Tile
//   P::num_tiles(prm), P::get_tile(prm, idx)            -> Tile (with .num_k and .n_cols)
//   P::load_begin(prm, tile) -> LoadCtx ; P::load(prm, tile, lctx, k, sA, sB, bar) -> TMA + expect_tx for chunk k
//     (the producer is ONE thread: anything but adds and a compare per chunk shows up as idle tensor pipes)
//   P::epilogue(prm, tile, row, col0, v[32], ncols)     -> consume 32 (or 16) accumulator columns
#pragma once
#include "tc_common.cuh"

namespace vsr {

constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = 128 * 128;  // 128 rows x 128 B

template <class P>
constexpr int tc_smem_bytes() {
  return P::STAGES * (TC_A_BYTES + P::BN * 128) + 1024;
}

// i-th tile of CTA (or CTA pair) `who` out of `n_who`: round robin, or a host-computed LPT list
// order[who*stride + i] (-1 terminated) when the tiles have very different costs.
__device__ __forceinline__ int tc_tile_round_robin(int who, int n_who, int i, int ntiles) {
  const int t = who + i * n_who;
  return t < ntiles ? t : -1;
}
__device__ __forceinline__ int tc_tile_listed(const int* order, int stride, int who, int i) {
  return i < stride ? order[(size_t)who * stride + i] : -1;
}

enum : uint32_t { ERR_PRODUCER = 0x100, ERR_MMA_FULL = 0x200, ERR_MMA_TEMPTY = 0x300, ERR_EPI = 0x400 };

template <class P>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ typename P::Params prm) {
  constexpr int BN = P::BN;
  constexpr int STAGES = P::STAGES;
  constexpr uint32_t STAGE_BYTES = TC_A_BYTES + BN * 128;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N");

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  // per-epilogue-warp 32x33 fp32 transpose scratch (policies that store row-scattered data coalesce through it)
  __shared__ __align__(16) float epi_scratch[P::EPI_SCRATCH ? 4 * 32 * 33 : 4];

  const long long t_kernel0 = TC_PROF_NOW();
  (void)t_kernel0;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
    P::prefetch(prm);
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int ntiles = P::num_tiles(prm);

  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      TC_PROF_DECL(w_empty);
      for (int it = 0;; ++it) {
        const int t = P::tile_at(prm, (int)blockIdx.x, (int)gridDim.x, it, ntiles);
        if (t < 0) break;
        const typename P::Tile tile = P::get_tile(prm, t);
        typename P::LoadCtx lc = P::load_begin(prm, tile);  // per-tile invariants: the k loop must stay division-free
        {
          const int tn = P::tile_at(prm, (int)blockIdx.x, (int)gridDim.x, it + 1, ntiles);
          if (tn >= 0) P::prefetch_tile(prm, P::get_tile(prm, tn));  // pull the next tile's activations into L2
        }
        for (int k = 0; k < tile.num_k; ++k) {
          TC_PROF_WAIT(w_empty, smem_u32(&bar_empty[stage]), phase ^ 1, ERR_PRODUCER | stage);
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          P::load(prm, tile, lc, k, sA, sA + TC_A_BYTES, smem_u32(&bar_full[stage]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      TC_PROF_ADD(P::PROF_ID, 0, w_empty);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      TC_PROF_DECL(w_full);
      TC_PROF_DECL(w_tempty);
      for (int it = 0;; ++it) {
        const int t = P::tile_at(prm, (int)blockIdx.x, (int)gridDim.x, it, ntiles);
        if (t < 0) break;
        const typename P::Tile tile = P::get_tile(prm, t);
        TC_PROF_WAIT(w_tempty, smem_u32(&bar_tempty[as]), aphase ^ 1, ERR_MMA_TEMPTY | as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        const uint32_t idesc = umma_idesc_f16(128, BN, 0, P::B_MN_MAJOR);
        for (int k = 0; k < tile.num_k; ++k) {
          TC_PROF_WAIT(w_full, smem_u32(&bar_full[stage]), phase, ERR_MMA_FULL | stage);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          const uint32_t sB = sA + TC_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 x UMMA_K(16) = 64
            const uint64_t adesc = umma_desc_sw128(sA + kk * 32, 16, 1024);
            const uint64_t bdesc = P::B_MN_MAJOR ? umma_desc_sw128(sB + kk * 2048, 8192, 1024)
                                                 : umma_desc_sw128(sB + kk * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (k | kk) != 0);
          }
          umma_commit(smem_u32(&bar_empty[stage]));  // frees the smem stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(&bar_tfull[as]));  // accumulator ready for the epilogue
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      TC_PROF_ADD(P::PROF_ID, 1, w_full);
      TC_PROF_ADD(P::PROF_ID, 2, w_tempty);
    }
    __syncwarp();
  } else {
    const uint32_t quarter = warp & 3;  // TMEM lane quarter this warp may read
    const uint32_t row = quarter * 32 + lane;
    uint32_t as = 0, aphase = 0;
    TC_PROF_DECL(w_tfull);
    TC_PROF_DECL(busy);
    for (int it = 0;; ++it) {
      const int t = P::tile_at(prm, (int)blockIdx.x, (int)gridDim.x, it, ntiles);
      if (t < 0) break;
      const typename P::Tile tile = P::get_tile(prm, t);
      TC_PROF_WAIT(w_tfull, smem_u32(&bar_tfull[as]), aphase, ERR_EPI | as);
      tc_fence_after();
      const long long e0 = TC_PROF_NOW();
      (void)e0;
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * BN;
      typename P::RowCtx ctx = P::row_begin(prm, tile, row);
      if constexpr (BN >= 32) {
        for (int c = 0; c < tile.n_cols; c += 32) {
          float v[32];
          tmem_ld32(taddr + c, v);
          P::epilogue(prm, tile, ctx, row, c, v, (P::EPI_SCRATCH ? epi_scratch + quarter * 32 * 33 : nullptr));
        }
      } else {
        float v[32];
        tmem_ld16(taddr, v);
        P::epilogue(prm, tile, ctx, row, 0, v, (P::EPI_SCRATCH ? epi_scratch + quarter * 32 * 33 : nullptr));
      }
      P::row_end(prm, tile, ctx, row);
      tc_fence_before();
      __syncwarp();
      TC_PROF_ACC(busy, e0);
      if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[as]));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (warp == 2 && lane == 0) {
      TC_PROF_ADD(P::PROF_ID, 3, w_tfull);
      TC_PROF_ADD(P::PROF_ID, 4, busy);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    TC_PROF_ADD(P::PROF_ID, 5, TC_PROF_NOW() - t_kernel0);
    TC_PROF_ADD(P::PROF_ID, 6, 1);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace vsr
