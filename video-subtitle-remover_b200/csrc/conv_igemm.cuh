// Implicit-GEMM convolution on tcgen05 for NHWC fp16 activations (stride 1, arbitrary tap offsets).
//   M = output pixels (tiles of tile_w x tile_h <= 128 pixels of one frame)
//   N = output channels (BN per tile), K = taps x Cin in chunks of 64 channels
// The A operand of tap (dy,dx) is the activation tensor shifted by (dy,dx): one 4-D TMA box
// {64 ch, tile_w, tile_h, 1 frame} whose out-of-range pixels are zero-filled by the TMA unit, which
// is exactly the zero padding of the reference convs (auto_sttn.py:76-95,156-164,215-218).
// Stride-2 convs are run as stride-1 2x2-tap convs over a space-to-depth input (see weights.cuh).
#pragma once
#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"

namespace vsr {

enum ConvFlags : int {
  CONV_LRELU = 1,      // LeakyReLU(0.2) after bias
  CONV_RESIDUAL = 2,   // act += res32 (fp32) before the stores; out32 (if set) gets the fp32 value, out16 its fp16 rounding
  CONV_S2D_STORE = 4,  // store out16 space-to-depth: [T, H/2, W/2, 4*Cout], channel = (y&1)*2*Cout + (x&1)*Cout + c
  CONV_RELU = 0x100,   // plain ReLU after bias (graph runtime)
  CONV_SCALED = 0x200, // out = acc * alpha + bias * bias_scale, fp16 overflow reported through *overflow (graph runtime)
  CONV_TMA_STORE = 0x400,  // fp16 output through a smem-staged TMA store (out_map), see conv_store_tma
  CONV_FINAL = 8,      // decoder.6: tanh -> (x+1)/2*255 -> trunc u8 -> first visit store / 0.5-0.5 blend into comps
};

struct ConvParams {
  CUtensorMap in_map;  // 4-D {C, W, H, T} fp16, box {64, tile_w, tile_h, 1}, SWIZZLE_128B
  CUtensorMap w_map;   // 2-D {K_total, Cout_pad} fp16, box {64, BN}, SWIZZLE_128B
  CUtensorMap w_map_half;  // same tensor, box {64, 128}: the half of B each CTA of a pair loads (Conv2Policy)
  CUtensorMap w_map_quarter;  // box {64, 64}: what each CTA of a 4-CTA cluster loads and multicasts (conv_halo.cuh, CL = 4)
  int T, H, W;         // output (= input) spatial size
  int tile_w, tile_h, tiles_x, tiles_y;
  int n_tiles;         // Cout_pad / BN
  int ntaps, cin_chunks;
  int cout;            // real output channels
  int flags;
  int prefetch;        // 1: L2-prefetch the next tile's activations (VSR_CONV_PREFETCH)
  int8_t tap_dy[81], tap_dx[81];  // up to 9x9 kernels
  const float* bias;   // [Cout_pad]
  __half* out16;       // NHWC fp16, pixel pitch out16_pitch (elements), channel offset out16_coff
  int out16_pitch, out16_coff;
  float* out32;        // NHWC fp32 residual stream [T,H,W,cout] (CONV_RESIDUAL)
  const float* res32;
  // CONV_FINAL
  float* comps;        // [chunk_T, H, W, 3] fp32 RGB running composites
  const int* frame_idx;   // window-local frame -> chunk frame
  const int* first_visit; // window-local frame -> 1 if this is the first time the frame is decoded
  // sttn-det low-res composite (sttn_det_inpaint.py:168): comp = mask>0 ? pred : input frame (RGB)
  const uint8_t* det_mask;  // [H,W] resized mask, nullptr for sttn-auto
  const uchar4* det_rgb;    // [chunk_T,H,W] RGBA8 input frames at model resolution
  // CONV_SCALED (graph runtime, rt_ops.cuh RtScale): out = acc * alpha + bias * bias_scale; *overflow = 1 when a
  // stored fp16 value leaves the format
  float alpha, bias_scale;
  int* overflow;
  // the output tensor is the window [crop_t, crop_t + out_H) x [crop_l, crop_l + out_W) of the computed grid (a conv over a
  // reflect-padded input stores only the interior); out_H = H, out_W = W, crop 0 for plain convs
  int crop_t, crop_l, out_H, out_W;
  // haloed-tile kernel (conv_halo.cuh): halo = max |tap offset|; in_map is then the box {64, 16, 16 + 2*halo, 1}, tile_w = 8 and
  // tile_h = the rows of the 16-row tile that are stored (tiles advance by tile_h); base_off = 0 leaves the descriptor's base-offset 0
  int halo, halo_base_off, halo_nb;
  // CONV_TMA_STORE: the fp16 output tensor as a 4-D map {cout, W, H, T} with box {32 ch, tile_w, tile_h, 1}, SWIZZLE_64B; scr_stride =
  // floats between the scratch areas of consecutive epilogue warps (the 4 warps of a column group pool theirs into two staging tiles)
  CUtensorMap out_map;
  int scr_stride;
};

// ---- TMA-store epilogue ------------------------------------------------------------------------------------------------------------
// The four epilogue warps that share a column range hold the 128 rows of the tile between them.  Per 32-column chunk each thread packs its
// row to fp16 (64 B) and writes it into a staging tile [128 rows][64 B] in the 64-byte swizzle TMA expects; one thread then issues ONE
// cp.async.bulk.tensor store of the box {32 ch, tile_w, tile_h} — the copy engine does the address generation, clips rows outside the
// tensor (partial tiles, the dummy half of an odd pair: frame index T) and writes whole lines.  Two staging tiles per group (8 KB each,
// carved from the warps' transposition scratch), so the store of chunk i overlaps the packing of chunk i + 1.  Measured motive: the 1x1
// Q/K/V projection executed 44 M warp instructions for 103 M outputs and held the tensor pipe at 18 % (profiles/ncu_r2c_gemm2.md).
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(src), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// v[32]: this thread's row (tile row `row`), channels [ch0, ch0 + 32) after bias / activation.  `parity` alternates per call (per group).
__device__ __forceinline__ void conv_store_tma(const ConvParams& p, int x0, int y0, int t, int row, int ch0, const float* v, float* scr, int& parity) {
  const int lane = threadIdx.x & 31, ew = (int)(threadIdx.x >> 5) - 2;
  const int wq = ew & 3, grp = ew >> 2;
  uint8_t* stage = reinterpret_cast<uint8_t*>(scr - wq * p.scr_stride) + parity * 8192;
  const bool leader = wq == 0 && lane == 0;
  if (leader) tma_store_wait_read<1>();          // the store that read this staging tile two chunks ago has finished reading it
  named_bar_sync(1 + grp, 128);
  uint4* dst = reinterpret_cast<uint4*>(stage + row * 64);
  const int sw = (row >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __align__(16) __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[8 * j + 2 * i], v[8 * j + 2 * i + 1]);
    dst[j ^ sw] = *reinterpret_cast<const uint4*>(h);
  }
  fence_proxy_async_smem();
  named_bar_sync(1 + grp, 128);
  if (leader) {
    tma_store_4d(&p.out_map, smem_u32(stage), ch0, x0, y0, t);
    tma_store_commit();
  }
  parity ^= 1;
}

// 32 accumulator columns [ch0, ch0 + 32) of the 32 pixel rows of one epilogue warp (row = lane), after bias / activation:
// optional fp32 residual add, fp32 store, fp16 store — every global access issued "transposed" (lane = 4 rows x 8 sixteen-byte
// chunks), data exchanged through the warp's 4 KB scratch with a 16-byte-chunk XOR swizzle (conflict-free both ways).
// pixv = output pixel index of this lane's row, or -1 when the row stores nothing.
__device__ __forceinline__ void conv_store_coalesced(const ConvParams& p, int pixv, int ch0, float* v, float* scr) {
  const int lane = threadIdx.x & 31;
  const int q = lane & 7, rsub = lane >> 3;
  float4* s4 = reinterpret_cast<float4*>(scr);
  const bool live = ch0 + 4 * q < p.cout;   // Cout padded up to the tile width: only the real channels are touched
  if ((p.flags & CONV_RESIDUAL) && p.out32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rr = 4 * j + rsub;
      const int pv = __shfl_sync(0xffffffffu, pixv, rr);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pv >= 0 && live) a = *reinterpret_cast<const float4*>(p.res32 + (size_t)pv * p.cout + ch0 + 4 * q);
      s4[rr * 8 + (q ^ (rr & 7))] = a;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = s4[lane * 8 + (i ^ (lane & 7))];
      v[4 * i] += a.x; v[4 * i + 1] += a.y; v[4 * i + 2] += a.z; v[4 * i + 3] += a.w;
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s4[lane * 8 + (i ^ (lane & 7))] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int rr = 4 * j + rsub;
    const int pv = __shfl_sync(0xffffffffu, pixv, rr);
    const float4 a = s4[rr * 8 + (q ^ (rr & 7))];
    if (pv >= 0 && live) {
      if (p.out32) *reinterpret_cast<float4*>(p.out32 + (size_t)pv * p.cout + ch0 + 4 * q) = a;
      if (p.out16) {
        __align__(8) __half2 h[2] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w)};
        *reinterpret_cast<uint2*>(p.out16 + (size_t)pv * p.out16_pitch + p.out16_coff + ch0 + 4 * q) = *reinterpret_cast<const uint2*>(h);
      }
    }
  }
  __syncwarp();
}

template <int BN_>
struct ConvPolicy {
  static constexpr int BN = BN_;
  static constexpr int STAGES = (BN_ == 256) ? 4 : 6;
  static constexpr int B_MN_MAJOR = 0;
  static constexpr int PROF_ID = (BN_ == 256) ? 0 : 4;
  static constexpr bool EPI_SCRATCH = false;
  using Params = ConvParams;
  struct Tile {
    int num_k, n_cols;
    int t, y0, x0, n0;
  };
  struct RowCtx {
    bool valid;
    int t, y, x;
    size_t pix;  // (t*H + y)*W + x
    int parity;  // CONV_TMA_STORE: which staging tile the next chunk uses (persists across tiles: the caller keeps it)
  };

  __device__ static void prefetch(const Params& p) {
    tma_prefetch_desc(&p.in_map);
    tma_prefetch_desc(&p.w_map);
  }
  __device__ static int num_tiles(const Params& p) { return p.T * p.tiles_y * p.tiles_x * p.n_tiles; }
  __device__ static int tile_at(const Params&, int who, int n_who, int i, int ntiles) { return tc_tile_round_robin(who, n_who, i, ntiles); }
  __device__ static Tile get_tile(const Params& p, int idx) {
    Tile t;
    const int n = idx % p.n_tiles;
    idx /= p.n_tiles;
    const int tx = idx % p.tiles_x;
    idx /= p.tiles_x;
    const int ty = idx % p.tiles_y;
    t.t = idx / p.tiles_y;
    t.x0 = tx * p.tile_w;
    t.y0 = ty * p.tile_h;
    t.n0 = n * BN;
    t.num_k = p.ntaps * p.cin_chunks;
    t.n_cols = BN;
    return t;
  }
  // first-touch activations of a tile: the un-shifted box of every 64-channel chunk (the taps re-read it from L2)
  __device__ static void prefetch_tile(const Params& p, const Tile& t) {
    if (!p.prefetch || t.t >= p.T) return;
    for (int kc = 0; kc < p.cin_chunks; ++kc) tma_prefetch_4d(&p.in_map, kc * 64, t.x0, t.y0, t.t);
  }
  struct LoadCtx {
    int tap, kc;        // running (tap, channel chunk) of the next k-chunk
    uint32_t tx_bytes;
  };
  __device__ static LoadCtx load_begin(const Params& p, const Tile&) {
    return LoadCtx{0, 0, (uint32_t)(p.tile_w * p.tile_h * 128 + BN * 128)};
  }
  __device__ static void load(const Params& p, const Tile& t, LoadCtx& lc, int k, uint32_t sA, uint32_t sB, uint32_t bar) {
    mbar_expect_tx(bar, lc.tx_bytes);
    tma_load_4d(sA, &p.in_map, bar, lc.kc * 64, t.x0 + p.tap_dx[lc.tap], t.y0 + p.tap_dy[lc.tap], t.t);
    tma_load_2d(sB, &p.w_map, bar, k * 64, t.n0);
    if (++lc.kc == p.cin_chunks) { lc.kc = 0; ++lc.tap; }
  }
  __device__ static RowCtx row_begin(const Params& p, const Tile& t, int row) {
    RowCtx c;
    const int ry = row / p.tile_w, rx = row - ry * p.tile_w;
    c.t = t.t;
    c.y = t.y0 + ry;
    c.x = t.x0 + rx;
    const int oy = c.y - p.crop_t, ox = c.x - p.crop_l;
    c.valid = (ry < p.tile_h) && (c.y < p.H) && (c.x < p.W) && ((unsigned)oy < (unsigned)p.out_H) && ((unsigned)ox < (unsigned)p.out_W);
    c.pix = ((size_t)t.t * p.out_H + oy) * p.out_W + ox;
    c.parity = 0;   // every tile stores an even number of chunks per column group, so the staging tiles keep alternating across tiles
    return c;
  }
  __device__ static void row_end(const Params&, const Tile&, RowCtx&, int) {}
  // `scr`: 4 KB of shared memory per epilogue warp, or nullptr.  With it, the fp32 residual read and the fp32 / fp16 stores go through a
  // 32 x 32 transposition so that every warp-wide memory instruction covers whole 128-byte lines (4 pixel rows x 128 B) instead of
  // 16 bytes of 32 different lines: the row-per-thread form costs 32 L1TEX wavefronts per instruction, which made the residual
  // epilogue (20 such instructions per 32 columns) as long as the tile's MMAs (ncu, profiles/ncu_r2a_conv_halo.md: tensor pipe 48 %).
  __device__ static void epilogue(const Params& p, const Tile& t, RowCtx& c, int row, int col0, float* v, float* scr) {
    if (scr == nullptr && !c.valid) return;
    constexpr int NV = (BN >= 32) ? 32 : 16;
    const int ch0 = t.n0 + col0;
    if (p.flags & CONV_SCALED) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(p.bias + ch0 + i);
        v[i] = fmaf(v[i], p.alpha, b.x * p.bias_scale);
        v[i + 1] = fmaf(v[i + 1], p.alpha, b.y * p.bias_scale);
        v[i + 2] = fmaf(v[i + 2], p.alpha, b.z * p.bias_scale);
        v[i + 3] = fmaf(v[i + 3], p.alpha, b.w * p.bias_scale);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        const float4 b = *reinterpret_cast<const float4*>(p.bias + ch0 + i);
        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
      }
    }
    if (p.flags & CONV_LRELU) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
    }
    if (p.flags & CONV_RELU) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if ((p.flags & CONV_SCALED) && p.overflow) {
      bool bad = false;
#pragma unroll
      for (int i = 0; i < NV; ++i) bad |= !(fabsf(v[i]) <= 65504.f);
      if (bad && c.valid) *p.overflow = 1;
    }
    if (p.flags & CONV_FINAL) {
      if (col0 != 0 || !c.valid) return;
      const int f = p.frame_idx[c.t];
      float* dst = p.comps + (((size_t)f * p.H + c.y) * p.W + c.x) * 3;
      const bool first = p.first_visit[c.t] != 0;
      bool keep_input = false;
      uchar4 in_px = make_uchar4(0, 0, 0, 0);
      if (p.det_mask) {
        keep_input = p.det_mask[(size_t)c.y * p.W + c.x] == 0;
        in_px = p.det_rgb[((size_t)f * p.H + c.y) * p.W + c.x];
      }
      const float in_rgb[3] = {(float)in_px.x, (float)in_px.y, (float)in_px.z};
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        // sttn_auto_inpaint.py:150-158: tanh -> (x+1)/2 -> *255 -> astype(uint8) (truncation)
        float y = (tanhf(v[i]) + 1.0f) * 0.5f;
        y = y * 255.0f;
        float q = (float)(unsigned char)fminf(fmaxf(y, 0.f), 255.f);
        if (keep_input) q = in_rgb[i];
        dst[i] = first ? q : (dst[i] * 0.5f + q * 0.5f);  // :159-162
      }
      return;
    }
    if (scr != nullptr && (p.flags & CONV_TMA_STORE)) {
      conv_store_tma(p, t.x0, t.y0, t.t, row, ch0, v, scr, c.parity);
      return;
    }
    if (scr != nullptr && !(p.flags & CONV_S2D_STORE)) {
      conv_store_coalesced(p, c.valid ? (int)c.pix : -1, ch0, v, scr);
      return;
    }
    if (!c.valid) return;
    if (p.out32) {
      float* o = p.out32 + c.pix * p.cout + ch0;
      if (p.flags & CONV_RESIDUAL) {
        const float* r = p.res32 + c.pix * p.cout + ch0;
#pragma unroll
        for (int i = 0; i < NV; i += 4) {
          const float4 a = *reinterpret_cast<const float4*>(r + i);
          v[i] += a.x; v[i + 1] += a.y; v[i + 2] += a.z; v[i + 3] += a.w;
        }
      }
#pragma unroll
      for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
    if (p.out16) {
      __half* o;
      if (p.flags & CONV_S2D_STORE) {
        const size_t opix = ((size_t)c.t * (p.H >> 1) + (c.y >> 1)) * (p.W >> 1) + (c.x >> 1);
        o = p.out16 + opix * p.out16_pitch + ((c.y & 1) * 2 + (c.x & 1)) * p.cout + ch0;
      } else {
        o = p.out16 + c.pix * p.out16_pitch + p.out16_coff + ch0;
      }
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        if (ch0 + i >= p.cout) break;  // Cout padded up to the tile width: only the real channels are stored
        __align__(16) __half2 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[i + 2 * j], v[i + 2 * j + 1]);
        *reinterpret_cast<uint4*>(o + i) = *reinterpret_cast<const uint4*>(h);
      }
    }
  }
};

// CTA-pair version of ConvPolicy<256>: pair-tile = two consecutive 128-pixel sub-tiles x 256 channels.
struct Conv2Policy {
  static constexpr int STAGES = 5;   // 5 x 32 KB + 34 KB of epilogue staging (6 stages would sit exactly on the 227 KB limit)
  static constexpr int PROF_ID = 1;
  static constexpr bool EPI_SCRATCH = true;
  static constexpr int EPI_WARPS = 8;   // two warps per TMEM lane quarter, each takes half of the 256 columns
  static constexpr int B_MN_MAJOR = 0;
  using Base = ConvPolicy<256>;
  using Params = ConvParams;
  using Tile = Base::Tile;
  using RowCtx = Base::RowCtx;
  __device__ static void prefetch(const Params& p) {
    tma_prefetch_desc(&p.in_map);
    tma_prefetch_desc(&p.w_map_half);
  }
  __device__ static int num_tiles(const Params& p) { return ((p.T * p.tiles_y * p.tiles_x + 1) >> 1) * p.n_tiles; }
  __device__ static int tile_at(const Params&, int who, int n_who, int i, int ntiles) { return tc_tile_round_robin(who, n_who, i, ntiles); }
  __device__ static Tile get_tile(const Params& p, int idx, uint32_t rank) {
    Tile t;
    const int n = idx % p.n_tiles;
    int sub = (idx / p.n_tiles) * 2 + (int)rank;
    const int total = p.T * p.tiles_y * p.tiles_x;
    t.n0 = n * 256;
    t.num_k = p.ntaps * p.cin_chunks;
    t.n_cols = 256;
    if (sub >= total) {  // odd tile count: the pair's second half is a dummy (TMA zero-fills frame index T)
      t.t = p.T; t.y0 = 0; t.x0 = 0;
      return t;
    }
    const int tx = sub % p.tiles_x;
    sub /= p.tiles_x;
    t.x0 = tx * p.tile_w;
    t.y0 = (sub % p.tiles_y) * p.tile_h;
    t.t = sub / p.tiles_y;
    return t;
  }
  __device__ static void prefetch_tile(const Params& p, const Tile& t) { Base::prefetch_tile(p, t); }
  using LoadCtx = Base::LoadCtx;
  __device__ static LoadCtx load_begin(const Params& p, const Tile&, uint32_t) {
    // the leader registers the bytes of both CTAs; the peer's loads may land first (tx-count is signed)
    return LoadCtx{0, 0, 2u * (uint32_t)(p.tile_w * p.tile_h * 128 + 128 * 128)};
  }
  __device__ static void load(const Params& p, const Tile& t, LoadCtx& lc, int k, uint32_t sA, uint32_t sB, uint32_t local_full,
                              uint32_t leader_full, uint32_t rank) {
    if (rank == 0) mbar_expect_tx(local_full, lc.tx_bytes);
    tma_load_4d_2sm(sA, &p.in_map, leader_full, lc.kc * 64, t.x0 + p.tap_dx[lc.tap], t.y0 + p.tap_dy[lc.tap], t.t);
    tma_load_2d_2sm(sB, &p.w_map_half, leader_full, k * 64, t.n0 + (int)rank * 128);
    if (++lc.kc == p.cin_chunks) { lc.kc = 0; ++lc.tap; }
  }
  __device__ static RowCtx row_begin(const Params& p, const Tile& t, int row) {
    RowCtx c = Base::row_begin(p, t, row);
    c.valid = c.valid && (t.t < p.T);
    return c;
  }
  __device__ static void row_end(const Params&, const Tile&, RowCtx&, int) {}
  __device__ static void epilogue(const Params& p, const Tile& t, RowCtx& c, int row, int col0, float* v, float* scr) {
    Base::epilogue(p, t, c, row, col0, v, scr);
  }
};

}  // namespace vsr
