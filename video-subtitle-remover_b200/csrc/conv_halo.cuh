// Haloed-tile implicit-GEMM convolution on CTA pairs (tcgen05 cta_group::2) for k x k stride-1 convs with Cout tiles of 256.
//
// Why: the per-tap kernel (conv_igemm.cuh, Conv2Policy) re-fetches the 128-pixel A box of its tile for every tap — 9 x 16 KB per
// 64-channel chunk for a 3x3 conv — next to 9 x 16 KB of weights, 64 B per SM and cycle at the tensor peak (8192 FLOP/cycle/SM):
// more than the ~43 B/cycle/SM the L2 delivers when every SM pulls (B300_MICROARCH.md: LTS cap ~6300 B/cycle chip-wide).  ncu showed
// it (round 1: l1tex__m_xbar2l1tex_read_bytes 1.5 GB per launch for 72 MB of algorithmic operands, tensor pipe 45 %).
//
// Here the A operand of ALL taps of a 64-channel chunk is ONE haloed tile in shared memory:
//   tile = 8 pixels wide x 16 rows (M = 128: UMMA row r = pixel (r / 8, r % 8)), halo h = dilation * (k / 2) <= 4;
//   one 4-D TMA box {64 ch, 16 px, 16 + 2h rows, 1 frame} at (x0 - h, y0 - h) per chunk (out-of-range pixels zero-filled: the conv
//   padding), written with SWIZZLE_128B at a row pitch of 16 pixel lines = 2048 B;
//   the A descriptor of tap (dy, dx) starts at line (dy + h) * 16 + (dx + h): its 8-row groups (one tile row each) are 2048 B apart
//   (SBO), and because 2048 is a multiple of the 1024-byte swizzle period every group sees the same swizzle phase, which the
//   descriptor's base-offset field ((start >> 7) & 7) carries.  No data is moved or re-fetched between taps.
// Operand traffic per 64-channel chunk and CTA drops from 9 x (16 + 16) KB to (36..40) + 9 x 16 KB: 36 B per SM and cycle at peak.
//
// Pipeline (per CTA; the leader = cluster rank 0 issues the MMAs of the pair, as in tc_gemm2.cuh):
//   warp 0 : TMA producer — A ring (NA haloed tiles), B ring (NB stages of this CTA's 128 of the 256 weight rows x 64 K)
//   warp 1 : MMA issuer (leader only): per chunk wait A, per tap wait B, 4 x tcgen05.mma (K = 16), commit B stage / A buffer
//   warps 2..5 : epilogue (ConvPolicy<256>::epilogue), accumulators double-buffered in TMEM
#pragma once
#include "conv_igemm.cuh"

namespace vsr {

constexpr int HALO_NA = 2;                  // A ring depth
constexpr int HALO_NB = 7;                  // B ring depth (at most; ConvParams::halo_nb stages are used)
constexpr int HALO_B_BYTES = 128 * 128;      // one B stage of the 256-wide kernel (this CTA's 128 weight rows x 64 K); BN / 2 rows in general
__host__ __device__ constexpr int halo_b_bytes(int bn) { return (bn / 2) * 128; }
constexpr int HALO_EPI_WARPS = 8;           // two per TMEM lane quarter, 128 columns each
constexpr int HALO_THREADS = 64 + 32 * HALO_EPI_WARPS;
constexpr int HALO_SCRATCH = HALO_EPI_WARPS * 4096;   // static: the epilogue's transposition scratch (conv_store_coalesced)
__host__ __device__ constexpr int halo_a_bytes(int halo) { return 16 * (16 + 2 * halo) * 128; }   // haloed rows x 16 lines x 128 B
// B stages that fit beside the two A buffers, the scratch and the barriers in 227 KB
inline int halo_b_stages(int halo, int bn = 256) {
  const int room = 232448 - HALO_SCRATCH - 512 - 1024 - HALO_NA * halo_a_bytes(halo);
  return room / halo_b_bytes(bn) < HALO_NB ? room / halo_b_bytes(bn) : HALO_NB;
}
inline int halo_smem_bytes(int halo, int bn = 256) { return HALO_NA * halo_a_bytes(halo) + halo_b_stages(halo, bn) * halo_b_bytes(bn) + 1024; }

enum : uint32_t { ERR_HALO_PROD_A = 0x500, ERR_HALO_PROD_B = 0x600, ERR_HALO_MMA_A = 0x700, ERR_HALO_MMA_B = 0x800,
                  ERR_HALO_MMA_T = 0x900, ERR_HALO_EPI = 0xA00 };

// SWIZZLE_128B K-major descriptor with a start address that is 128-byte- but not 1024-byte-aligned
__device__ __forceinline__ uint64_t umma_desc_sw128_off(uint32_t saddr, uint32_t sbo_bytes, int with_base_offset) {
  uint64_t d = umma_desc_sw128(saddr, 16, sbo_bytes);
  if (with_base_offset) d |= (uint64_t)((saddr >> 7) & 7u) << 49;   // base offset: phase of the start line inside the 8-line swizzle period
  return d;
}

// TMA load of a 2-D box into the same smem offset of every CTA in `cta_mask` (multicast through L2); with cta_group::2 the completion
// of each destination CTA is signalled on the barrier at `bar`'s offset in that CTA or in its pair peer, exactly as the issuing CTA
// addresses it (here: always the pair's even CTA, the MMA leader).
__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(dst),
      "l"(m), "r"(cluster_bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}

// CL = 2: one CTA pair per cluster.  CL = 4: two pairs per cluster work on neighbouring pixel tiles of the same Cout tile and share the
// weights: every CTA fetches a quarter of the 256 x 64 weight chunk (64 rows) and multicasts it to the CTA of the other pair that
// needs the same half, so the L2 -> SM weight traffic per CTA halves (16 -> 8 KB per chunk and tap).
// BN = Cout tile of the pair (256, 128 or 64): each CTA holds BN / 2 weight rows per stage; the narrow tiles serve the decoder convs
// (256 -> 128, 128 -> 64, 64 -> 64), whose per-tap kernels re-fetched a 16 KB A box for as little as 1 MFLOP.
template <int CL, int BN>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(HALO_THREADS, 1) conv_halo_kernel(const __grid_constant__ ConvParams prm) {
  static_assert(CL == 2 || CL == 4, "cluster of one or two CTA pairs");
  static_assert(BN == 256 || BN == 128 || BN == 64, "Cout tile");
  static_assert(CL == 2 || BN == 256, "the multicast variant is built for 256-wide tiles");
  using Base = ConvPolicy<BN>;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr uint32_t B_BYTES = (uint32_t)halo_b_bytes(BN);
  constexpr int NPAIR = CL / 2;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[HALO_NA], a_empty[HALO_NA], b_full[HALO_NB], b_empty[HALO_NB], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(1024) float epi_scratch[HALO_SCRATCH / 4];

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1;          // rank inside the pair (0 = MMA leader)
  const uint32_t pair = crank >> 1;         // pair inside the cluster
  const uint32_t leader = crank & ~1u;      // cluster rank of this pair's leader
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
  const uint16_t all_mask = (uint16_t)((1u << CL) - 1);
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int halo = prm.halo;
  const uint32_t a_bytes = (uint32_t)halo_a_bytes(halo);
  const uint32_t NB = (uint32_t)prm.halo_nb;
  const uint32_t sA0 = smem_base, sB0 = smem_base + HALO_NA * a_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < HALO_NA; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), 1);
    }
    for (int s = 0; s < HALO_NB; ++s) {
      mbar_init(smem_u32(&b_full[s]), 1);
      mbar_init(smem_u32(&b_empty[s]), NPAIR);   // a stage is free when every pair that receives the multicast has consumed it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bar_tfull[s]), 1);
      mbar_init(smem_u32(&bar_tempty[s]), 2 * HALO_EPI_WARPS);
    }
    fence_barrier_init();
    tma_prefetch_desc(&prm.in_map);
    tma_prefetch_desc(CL == 4 ? &prm.w_map_quarter : &prm.w_map_half);
  }
  if (warp == 1) tmem_alloc_2cta(smem_u32(&tmem_slot), TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // work list: groups of NPAIR neighbouring pair-tiles x Cout tiles; cluster c takes groups c, c + nclusters, ...
  const int sub_tiles = prm.T * prm.tiles_y * prm.tiles_x;
  const int pair_tiles_m = (sub_tiles + 1) >> 1;
  const int ngroups = ((pair_tiles_m + NPAIR - 1) / NPAIR) * prm.n_tiles;
  const int first = (int)cluster_id_x(), step = (int)cluster_count_x();
  auto tile_of = [&](int g) {   // this CTA's 128-pixel tile of group g (a dummy beyond the end: TMA zero-fills frame index T)
    typename Base::Tile t;
    const int n = g % prm.n_tiles;
    int sub = ((g / prm.n_tiles) * NPAIR + (int)pair) * 2 + (int)rank;
    t.n0 = n * BN;
    t.num_k = prm.ntaps * prm.cin_chunks;
    t.n_cols = BN;
    if (sub >= sub_tiles) {
      t.t = prm.T; t.y0 = 0; t.x0 = 0;
      return t;
    }
    t.x0 = (sub % prm.tiles_x) * prm.tile_w;
    sub /= prm.tiles_x;
    t.y0 = (sub % prm.tiles_y) * prm.tile_h;
    t.t = sub / prm.tiles_y;
    return t;
  };

  if (warp == 0) {
    if (elect_one()) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0;
      for (int g = first; g < ngroups; g += step) {
        const typename Base::Tile tile = tile_of(g);
        for (int kc = 0; kc < prm.cin_chunks; ++kc) {
          mbar_wait(smem_u32(&a_empty[as]), aph ^ 1, ERR_HALO_PROD_A | as);
          if (rank == 0) mbar_expect_tx(smem_u32(&a_full[as]), 2u * a_bytes);
          tma_load_4d_2sm(sA0 + as * a_bytes, &prm.in_map, mapa_cluster(smem_u32(&a_full[as]), leader), kc * 64, tile.x0 - halo,
                          tile.y0 - halo, tile.t);
          if (++as == HALO_NA) { as = 0; aph ^= 1; }
          for (int tap = 0; tap < prm.ntaps; ++tap) {
            mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1, ERR_HALO_PROD_B | bs);
            if (rank == 0) mbar_expect_tx(smem_u32(&b_full[bs]), 2u * B_BYTES);
            const uint32_t full = mapa_cluster(smem_u32(&b_full[bs]), leader);
            const int k0 = (tap * prm.cin_chunks + kc) * 64;
            if (CL == 4) {   // rows [pair * 64, +64) of this CTA's half, to the CTAs of both pairs with the same in-pair rank
              tma_load_2d_2sm_mc(sB0 + bs * B_BYTES + pair * (B_BYTES / 2), &prm.w_map_quarter, full, k0,
                                 tile.n0 + (int)rank * 128 + (int)pair * 64, (uint16_t)(5u << rank));
            } else {
              tma_load_2d_2sm(sB0 + bs * B_BYTES, &prm.w_map_half, full, k0, tile.n0 + (int)rank * (BN / 2));
            }
            if (++bs == NB) { bs = 0; bph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      uint32_t as = 0, aph = 0, bs = 0, bph = 0, acc = 0, accph = 0;
      const uint32_t idesc = umma_idesc_f16(256, BN, 0, 0);
      for (int g = first; g < ngroups; g += step) {
        mbar_wait(smem_u32(&bar_tempty[acc]), accph ^ 1, ERR_HALO_MMA_T | acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kc = 0; kc < prm.cin_chunks; ++kc) {
          mbar_wait(smem_u32(&a_full[as]), aph, ERR_HALO_MMA_A | as);
          tc_fence_after();
          const uint32_t sA = sA0 + as * a_bytes;
          for (int tap = 0; tap < prm.ntaps; ++tap) {
            mbar_wait(smem_u32(&b_full[bs]), bph, ERR_HALO_MMA_B | bs);
            tc_fence_after();
            const uint32_t sB = sB0 + bs * B_BYTES;
            const uint32_t a_tap = sA + (uint32_t)(((int)prm.tap_dy[tap] + halo) * 16 + ((int)prm.tap_dx[tap] + halo)) * 128u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16_2cta(d_tmem, umma_desc_sw128_off(a_tap + kk * 32, 2048, prm.halo_base_off), umma_desc_sw128(sB + kk * 32, 16, 1024), idesc,
                            (kc | tap | kk) != 0);
            umma_commit_2cta(smem_u32(&b_empty[bs]), all_mask);
            if (++bs == NB) { bs = 0; bph ^= 1; }
          }
          umma_commit_2cta(smem_u32(&a_empty[as]), pair_mask);
          if (++as == HALO_NA) { as = 0; aph ^= 1; }
        }
        umma_commit_2cta(smem_u32(&bar_tfull[acc]), pair_mask);
        if (++acc == 2) { acc = 0; accph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const uint32_t quarter = warp & 3;
    const uint32_t ew = warp - 2;
    const uint32_t row = quarter * 32 + lane;
    float* scr = epi_scratch + ew * 1024;
    const int c_begin = (int)(ew >> 2) * (BN / 2);
    uint32_t acc = 0, accph = 0;
    for (int g = first; g < ngroups; g += step) {
      const typename Base::Tile tile = tile_of(g);
      typename Base::RowCtx ctx = Base::row_begin(prm, tile, row);
      ctx.valid = ctx.valid && tile.t < prm.T;   // dummy half of an odd pair
      mbar_wait(smem_u32(&bar_tfull[acc]), accph, ERR_HALO_EPI | acc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * BN;
      for (int c = c_begin; c < c_begin + BN / 2; c += 32) {
        float v[32];
        tmem_ld32(taddr + c, v);
        Base::epilogue(prm, tile, ctx, row, c, v, scr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_cluster(smem_u32(&bar_tempty[acc]), leader));
      if (++acc == 2) { acc = 0; accph ^= 1; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the epilogue's TMA stores read this CTA's smem: finished before exit
  }

  tc_fence_before();
  cluster_sync_all();  // the peers may still be signalling our barriers / writing our smem until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

}  // namespace vsr
