// CTA-pair variant of the skeleton in tc_gemm.cuh (tcgen05 cta_group::2): two CTAs of a cluster (= two
// SMs of a TPC) own one 256 x 256 tile.  Each CTA TMA-loads its own 128 rows of A and its own half
// (128 rows) of B, so the L2->SM operand traffic per FLOP drops to 2/3 of the single-CTA 128x256
// tile — the single-CTA conv kernel was measured L2-fill-bound (1.06 GB of operands in 82 us).
// The leader CTA (cluster rank 0) issues the MMAs for the pair; every CTA keeps 128 rows of the
// accumulator in its own TMEM and runs its own epilogue.
//   full[s]    : in the leader, 1 arrival (the leader's expect_tx for the bytes of BOTH CTAs); the peer's TMA
//                loads complete_tx on it directly (no remote arrive on the critical path)
//   empty[s]   : per CTA, released by a multicast tcgen05.commit from the leader
//   tfull[a]   : per CTA, multicast commit;   tempty[a] : in the leader, 8 arrivals (4 warps x 2 CTAs)
#pragma once
#include "tc_gemm.cuh"

namespace vsr {

template <class P>
constexpr int tc2_smem_bytes() {
  return P::STAGES * (2 * TC_A_BYTES) + 1024;
}

// Epilogue warps per CTA: 4 (one per TMEM lane quarter) unless the policy asks for 8 (`EPI_WARPS = 8`: two warps per quarter, each
// taking half of the tile's columns — for policies whose column groups are independent and whose epilogue, not the MMA, is the
// longer of the two per tile: short-K GEMMs such as the 1x1 Q/K/V projection).
template <class P, class = void>
struct tc_epi_warps { static constexpr int value = 4; };
template <class P>
struct tc_epi_warps<P, decltype((void)P::EPI_WARPS)> { static constexpr int value = P::EPI_WARPS; };
template <class P>
constexpr int tc2_threads() { return 64 + 32 * tc_epi_warps<P>::value; }
// per-tile instruction-descriptor bits (operand formats) for policies that define idesc_extra(prm, tile)
template <class P, class = void>
struct tc_idesc_extra {
  __device__ static uint32_t get(const typename P::Params&, const typename P::Tile&) { return 0u; }
};
template <class P>
struct tc_idesc_extra<P, decltype((void)&P::idesc_extra)> {
  __device__ static uint32_t get(const typename P::Params& p, const typename P::Tile& t) { return P::idesc_extra(p, t); }
};

template <class P>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc2_threads<P>(), 1)
    tc_gemm2_kernel(const __grid_constant__ typename P::Params prm) {
  constexpr int EW = tc_epi_warps<P>::value;
  static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
  constexpr int BN = 256;
  constexpr int STAGES = P::STAGES;
  constexpr uint32_t STAGE_BYTES = 2 * TC_A_BYTES;  // A: 128 rows, B: this CTA's 128 of the 256 rows
  constexpr uint32_t TMEM_COLS = 512;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  // per-epilogue-warp 32x33 fp32 transpose scratch (policies that store row-scattered data coalesce through it)
  __shared__ __align__(1024) float epi_scratch[P::EPI_SCRATCH ? EW * 32 * 33 : 4];

  const long long t_kernel0 = TC_PROF_NOW();
  (void)t_kernel0;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 2 * EW);
    }
    fence_barrier_init();
    P::prefetch(prm);
  }
  if (warp == 1) tmem_alloc_2cta(smem_u32(&tmem_slot), TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const int ntiles = P::num_tiles(prm);
  const int first = (int)cluster_id_x(), step = (int)cluster_count_x();

  if (warp == 0) {
    if (elect_one()) {
      uint32_t stage = 0, phase = 0;
      TC_PROF_DECL(w_empty);
      for (int it = 0;; ++it) {
        const int t = P::tile_at(prm, first, step, it, ntiles);
        if (t < 0) break;
        const typename P::Tile tile = P::get_tile(prm, t, rank);
        typename P::LoadCtx lc = P::load_begin(prm, tile, rank);
        {
          const int tn = P::tile_at(prm, first, step, it + 1, ntiles);
          if (tn >= 0) P::prefetch_tile(prm, P::get_tile(prm, tn, rank));  // pull the next tile's activations into L2
        }
        for (int k = 0; k < tile.num_k; ++k) {
          TC_PROF_WAIT(w_empty, smem_u32(&bar_empty[stage]), phase ^ 1, ERR_PRODUCER | stage);
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          const uint32_t leader_full = mapa_cluster(smem_u32(&bar_full[stage]), 0);
          P::load(prm, tile, lc, k, sA, sA + TC_A_BYTES, smem_u32(&bar_full[stage]), leader_full, rank);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      TC_PROF_ADD(P::PROF_ID, 0, w_empty);
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      TC_PROF_DECL(w_full);
      TC_PROF_DECL(w_tempty);
      const uint32_t idesc0 = umma_idesc_f16(256, BN, 0, P::B_MN_MAJOR);
      for (int it = 0;; ++it) {
        const int t = P::tile_at(prm, first, step, it, ntiles);
        if (t < 0) break;
        const typename P::Tile tile = P::get_tile(prm, t, 0);
        const uint32_t idesc = idesc0 | tc_idesc_extra<P>::get(prm, tile);
        TC_PROF_WAIT(w_tempty, smem_u32(&bar_tempty[as]), aphase ^ 1, ERR_MMA_TEMPTY | as);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int k = 0; k < tile.num_k; ++k) {
          TC_PROF_WAIT(w_full, smem_u32(&bar_full[stage]), phase, ERR_MMA_FULL | stage);
          tc_fence_after();
          const uint32_t sA = smem_base + stage * STAGE_BYTES;
          const uint32_t sB = sA + TC_A_BYTES;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t bdesc = P::B_MN_MAJOR ? umma_desc_sw128(sB + kk * 2048, 8192, 1024)
                                                 : umma_desc_sw128(sB + kk * 32, 16, 1024);
            umma_f16_2cta(d_tmem, umma_desc_sw128(sA + kk * 32, 16, 1024), bdesc, idesc, (k | kk) != 0);
          }
          umma_commit_2cta(smem_u32(&bar_empty[stage]), 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(smem_u32(&bar_tfull[as]), 3);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      TC_PROF_ADD(P::PROF_ID, 1, w_full);
      TC_PROF_ADD(P::PROF_ID, 2, w_tempty);
    }
    __syncwarp();
  } else {
    const uint32_t quarter = warp & 3;
    const uint32_t ew = warp - 2;             // epilogue warp 0 .. EW-1
    const uint32_t row = quarter * 32 + lane;
    uint32_t as = 0, aphase = 0;
    TC_PROF_DECL(w_tfull);
    TC_PROF_DECL(busy);
    for (int it = 0;; ++it) {
      const int t = P::tile_at(prm, first, step, it, ntiles);
      if (t < 0) break;
      const typename P::Tile tile = P::get_tile(prm, t, rank);
      TC_PROF_WAIT(w_tfull, smem_u32(&bar_tfull[as]), aphase, ERR_EPI | as);
      tc_fence_after();
      const long long e0 = TC_PROF_NOW();
      (void)e0;
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + as * BN;
      typename P::RowCtx ctx = P::row_begin(prm, tile, row);
      const int c_begin = (EW == 8) ? (int)(ew >> 2) * (tile.n_cols >> 1) : 0;
      const int c_end = (EW == 8) ? c_begin + (tile.n_cols >> 1) : tile.n_cols;
      for (int c = c_begin; c < c_end; c += 32) {
        float v[32];
        tmem_ld32(taddr + c, v);
        P::epilogue(prm, tile, ctx, row, c, v, P::EPI_SCRATCH ? epi_scratch + ew * 32 * 33 : nullptr);
      }
      P::row_end(prm, tile, ctx, row);
      tc_fence_before();
      __syncwarp();
      TC_PROF_ACC(busy, e0);
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&bar_tempty[as]), 0));
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // TMA stores of a policy's epilogue read this CTA's smem: finished before exit
    if (warp == 2 && lane == 0) {
      TC_PROF_ADD(P::PROF_ID, 3, w_tfull);
      TC_PROF_ADD(P::PROF_ID, 4, busy);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be signalling our barriers / reading our smem until here
  if (threadIdx.x == 0) {
    TC_PROF_ADD(P::PROF_ID, 5, TC_PROF_NOW() - t_kernel0);
    TC_PROF_ADD(P::PROF_ID, 6, 1);
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

}  // namespace vsr
