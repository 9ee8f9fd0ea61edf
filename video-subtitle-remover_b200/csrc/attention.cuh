// Multi-geometry patch attention of STTN (auto_sttn.py:140-206) on tcgen05, reading Q/K/V straight
// out of the NHWC [T,H,W,3*256] fp16 projection buffer — the reference's view/permute/contiguous
// soft-split copies (auto_sttn.py:182-190, 201-202) do not exist here.
//
// Head i owns channels [64 i, 64 i + 64) and patch (pw, ph): token = (t, oh, ow), feature = the
// ph*pw pixels of the patch x 64 channels.  A dot product over the feature axis is order-invariant,
// so K-chunk `pos = py*pw + px` of a token is simply the 64 contiguous channels of patch pixel
// (py, px): a 5-D TMA view {64 ch, px, ow, py, t*oh} of the buffer turns "pixel (py,px) of 128 tokens"
// into one box {64,1,owp,1,128/owp}.  owp = ow rounded up to a power of two; the padded token index is
// kp = toh*owp + owi (pad slots are zero-filled by TMA and masked in the softmax).
//
//   Big problems (many tiles, `fused`): the scores never leave the SM.  Pass A computes S tiles and keeps
//   only the per-(row, key-tile) max; pass B recomputes S and writes P = exp((S - rowmax)/sqrt(D)) in fp16 plus
//   per-(row, key-tile) partial row sums (plain stores, fixed summation order -> deterministic).  Recomputing
//   Q.K^T costs 2/3 more tensor work than storing S, but the fp32 S round trip (write + two reads + a
//   softmax kernel) was measured to cost more than twice that.
//   Small problems (few tiles but a huge feature dim, e.g. 60 tokens x 76 800): split-K, each split writes
//   its own fp32 slab of S; softmax_rows_kernel adds the slabs in a fixed order and writes P / rowsum.
//   pv     : O[qp, pos,:] = (sum_kp P[qp,kp] V_pos[kp,:]) / rowsum[qp]  -> written back in NHWC
#pragma once
#include <cuda_bf16.h>

#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"

namespace vsr {

struct AttnHead {
  int pw, ph, ow, oh, owp;
  int npos;        // ph*pw K-chunks of the score GEMM / 64-wide column groups of the PV GEMM
  int toh_total;   // T*oh
  int ntt;         // token tiles of 128 padded tokens
  int nk64;        // 64-token K-chunks of the PV GEMM that contain real tokens
  int splits, chunks_per_split;  // score GEMM split-K
  int score_work_begin, pv_work_begin;
  int score_work_begin2, pv_work_begin2;  // CTA-pair work lists
  int pv_ntiles;   // ceil(npos / 4)
  int ldS, ldP;    // row pitches (elements)
  long long slabS; // elements between split-K slabs of S (ntt*128*ldS)
  float scale_log2e;  // log2(e) / sqrt(64*ph*pw)
  float* S;
  __half* P;
  float* rowsum;
  int fused;            // 0: split-K slabs of S + softmax kernel; 1: two-pass score kernels (rowmax_part / rowsum_part);
                        // 2: "direct" — ONE score pass that writes P = exp2((s - s_self) * log2e / sqrt(D)) + per-tile row sums, where the
                        //    row shift s_self = q_i . k_i is the token's logit with itself (attn_diag_kernel; softmax is shift invariant):
                        //    the row maximum is >= it, so every row holds a 1 and nothing under-flows; a key that beats "self" by more
                        //    than 2^15.5 raises the range flag.  No S, no row-max pass, no softmax kernel
  int npairs;           // key tile pairs = ceil(ntt / 2)
  float* rowmax_part;   // [ntt*128][npairs]
  float* rowsum_part;   // [ntt*128][npairs]
  float* rowshift;      // [ntt*128] direct heads: q_i . k_i (un-scaled), 0 for pad rows
  int scoreB_begin, scoreB_begin2;  // pass-B work lists (fused problems only)
};

// 2^x on the SFU (ex2.approx: 2 ulp, exact 0 for -inf) — the softmax is bandwidth-bound only if its math stays cheap
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int ATTN_MAX_HEADS = 8;  // (windows in a group) x (patch geometries)

struct ScoreParams {
  CUtensorMap qmap[ATTN_MAX_HEADS], kmap[ATTN_MAX_HEADS];
  AttnHead h[ATTN_MAX_HEADS];
  int nheads, total_work, total_work2;
  int pass;  // 0 = pass A (max / S slabs), 1 = pass B (P)
  int totalB, totalB2;
  // LPT schedule of pass A (tiles cost between 1 and >100 K-chunks): order[who*order_stride + i], -1 terminated;
  // nullptr = round robin
  const int* order;
  int order_stride;
  int softmax_row_begin[ATTN_MAX_HEADS + 1];  // first softmax block of each problem (fused problems: empty range)
  int* overflow;   // direct heads: set when an exponent leaves the range fp16 holds without a row shift (the host then re-runs unfused)
};

struct ScorePolicy {
  static constexpr int BN = 256;  // two 128-token key tiles per CTA tile: 96 B/clk of operand fill instead of 128
  static constexpr int STAGES = 4;
  static constexpr int B_MN_MAJOR = 0;
  static constexpr int PROF_ID = 2;
  static constexpr bool EPI_SCRATCH = true;
  using Params = ScoreParams;
  struct Tile {
    int num_k, n_cols;
    int head, qi, kj, kbeg, split, nkt;  // nkt = key tiles (1 or 2) in this CTA tile
  };
  struct RowCtx {
    float* dst;
    float m, sum;  // pass A: running max of this row over the tile; pass B: row max, running sum
  };
  __device__ static void prefetch(const Params& p) {
    for (int i = 0; i < p.nheads; ++i) {
      tma_prefetch_desc(&p.qmap[i]);
      tma_prefetch_desc(&p.kmap[i]);
    }
  }
  __device__ static int num_tiles(const Params& p) { return p.pass ? p.totalB : p.total_work; }
  __device__ static int tile_at(const Params& p, int who, int n_who, int i, int ntiles) {
    return (p.order && p.pass == 0) ? tc_tile_listed(p.order, p.order_stride, who, i) : tc_tile_round_robin(who, n_who, i, ntiles);
  }
  __device__ static Tile get_tile(const Params& p, int idx) {
    int hd = 0;
    if (p.pass) {
      while (hd + 1 < p.nheads && idx >= p.h[hd + 1].scoreB_begin) ++hd;
      idx -= p.h[hd].scoreB_begin;
    } else {
      while (hd + 1 < p.nheads && idx >= p.h[hd + 1].score_work_begin) ++hd;
      idx -= p.h[hd].score_work_begin;
    }
    const AttnHead& h = p.h[hd];
    Tile t;
    t.head = hd;
    const int sp = idx % h.splits;
    idx /= h.splits;
    const int nkt2 = (h.ntt + 1) >> 1;
    t.kj = idx % nkt2;
    t.qi = idx / nkt2;
    t.kbeg = sp * h.chunks_per_split;
    t.num_k = min(h.chunks_per_split, h.npos - t.kbeg);
    t.nkt = min(2, h.ntt - 2 * t.kj);
    t.n_cols = 128 * t.nkt;
    t.split = sp;
    return t;
  }
  __device__ static void prefetch_tile(const Params&, const Tile&) {}
  struct LoadCtx {
    int px, py, pw;          // patch pixel of the next k-chunk (advanced without divisions)
    int qrow, krow0, krow1;  // 5th TMA coordinate of the query tile and of the key tile(s)
    uint32_t tx_bytes;
  };
  __device__ static LoadCtx load_begin(const Params& p, const Tile& t) {
    const AttnHead& h = p.h[t.head];
    const int rows = 128 / h.owp;
    LoadCtx lc;
    lc.pw = h.pw;
    lc.py = t.kbeg / h.pw;
    lc.px = t.kbeg - lc.py * h.pw;
    lc.qrow = t.qi * rows;
    lc.krow0 = (2 * t.kj) * rows;
    lc.krow1 = (2 * t.kj + 1) * rows;
    lc.tx_bytes = (uint32_t)((1 + t.nkt) * TC_A_BYTES);
    return lc;
  }
  __device__ static void load(const Params& p, const Tile& t, LoadCtx& lc, int, uint32_t sA, uint32_t sB, uint32_t bar) {
    mbar_expect_tx(bar, lc.tx_bytes);
    tma_load_5d(sA, &p.qmap[t.head], bar, 0, lc.px, 0, lc.py, lc.qrow);
    tma_load_5d(sB, &p.kmap[t.head], bar, 0, lc.px, 0, lc.py, lc.krow0);
    if (t.nkt == 2) tma_load_5d(sB + TC_A_BYTES, &p.kmap[t.head], bar, 0, lc.px, 0, lc.py, lc.krow1);
    if (++lc.px == lc.pw) { lc.px = 0; ++lc.py; }
  }
  __device__ static RowCtx row_begin(const Params& p, const Tile& t, int row) {
    const AttnHead& h = p.h[t.head];
    RowCtx c;
    c.dst = h.S + (size_t)t.split * h.slabS + (size_t)(t.qi * 128 + row) * h.ldS + t.kj * 256;
    c.m = -INFINITY;
    c.sum = 0.f;
    if (h.fused == 2) {
      c.m = t.n_cols > 0 ? h.rowshift[t.qi * 128 + row] : 0.f;   // direct: shift by the token's own logit
    } else if (h.fused && p.pass == 1 && t.n_cols > 0) {  // row max = max over the key-tile partials of pass A
      const float* pm = h.rowmax_part + (size_t)(t.qi * 128 + row) * h.npairs;
      float m = pm[0];
      for (int j = 1; j < h.npairs; ++j) m = fmaxf(m, pm[j]);
      c.m = m;
    }
    return c;
  }
  // The accumulator arrives one row per thread; storing it like that scatters every 16-byte store over 32
  // different rows (measured: the epilogue was 84 % of this kernel).  Transpose 32x32 blocks through smem so
  // that each store instruction writes whole 64/128-byte row segments.
  __device__ static void epilogue(const Params& p, const Tile& t, RowCtx& c, int row, int col0, float* v, float* scr) {
    const AttnHead& h = p.h[t.head];
    const int lane = row & 31;
    if (!h.fused) {  // split-K problem: fp32 slab of S
#pragma unroll
      for (int i = 0; i < 32; ++i) scr[lane * 33 + i] = v[i];
      __syncwarp();
      float* base = c.dst - (size_t)lane * h.ldS + col0;  // row 0 of this warp's 32 rows
      const int r4 = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int rr = j * 4 + r4;
        const float* sp = scr + rr * 33 + c4;
        *reinterpret_cast<float4*>(base + (size_t)rr * h.ldS + c4) = make_float4(sp[0], sp[1], sp[2], sp[3]);
      }
      __syncwarp();
      return;
    }
    const int kp0 = t.kj * 256 + col0;
    const int nvalid = h.toh_total * h.owp, owm = h.owp - 1;
    if (p.pass == 0) {  // pass A: only the max of the valid columns survives
      float m = c.m;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kp = kp0 + i;
        if (kp < nvalid && (kp & owm) < h.ow) m = fmaxf(m, v[i]);
      }
      c.m = m;
      return;
    }
    // pass B: P = exp2((s - max) * log2(e)/sqrt(D)) in fp16 (direct heads: the token's own logit instead of the max), masked slots = 0;
    // row sums of the rounded values
    uint32_t* scw = reinterpret_cast<uint32_t*>(scr);
    float sum = c.sum;
    if (h.fused == 2) {
      bool big = false;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int kp = kp0 + i;
        const float a0 = (v[i] - c.m) * h.scale_log2e, a1 = (v[i + 1] - c.m) * h.scale_log2e;
        const bool ok0 = kp < nvalid && (kp & owm) < h.ow, ok1 = kp + 1 < nvalid && ((kp + 1) & owm) < h.ow;
        big |= (ok0 && a0 > 15.5f) | (ok1 && a1 > 15.5f);
        const float e0 = ok0 ? fast_exp2(fminf(a0, 15.9f)) : 0.f;
        const float e1 = ok1 ? fast_exp2(fminf(a1, 15.9f)) : 0.f;
        const __half2 hh = __floats2half2_rn(e0, e1);
        const float2 f = __half22float2(hh);
        sum += f.x + f.y;
        scw[lane * 17 + (i >> 1)] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      if (big && p.overflow) *p.overflow = 1;
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int kp = kp0 + i;
        const float e0 = (kp < nvalid && (kp & owm) < h.ow) ? exp2f((v[i] - c.m) * h.scale_log2e) : 0.f;
        const float e1 = (kp + 1 < nvalid && ((kp + 1) & owm) < h.ow) ? exp2f((v[i + 1] - c.m) * h.scale_log2e) : 0.f;
        const __half2 hh = __floats2half2_rn(e0, e1);
        const float2 f = __half22float2(hh);
        sum += f.x + f.y;
        scw[lane * 17 + (i >> 1)] = *reinterpret_cast<const uint32_t*>(&hh);
      }
    }
    c.sum = sum;
    __syncwarp();
    // 32 rows x 64 bytes: each store instruction writes 8 rows x 64 contiguous bytes
    __half* pbase = h.P + (size_t)(t.qi * 128 + (row & ~31)) * h.ldP + kp0;
    const int r8 = lane >> 2, c4 = (lane & 3) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = j * 8 + r8;
      const uint32_t* sp = scw + rr * 17 + c4;
      *reinterpret_cast<uint4*>(pbase + (size_t)rr * h.ldP + c4 * 2) = make_uint4(sp[0], sp[1], sp[2], sp[3]);
    }
    __syncwarp();
  }
  __device__ static void row_end(const Params& p, const Tile& t, RowCtx& c, int row) {
    const AttnHead& h = p.h[t.head];
    if (!h.fused || t.n_cols == 0) return;
    const size_t idx = (size_t)(t.qi * 128 + row) * h.npairs + t.kj;
    if (p.pass == 0) h.rowmax_part[idx] = c.m;
    else h.rowsum_part[idx] = c.sum;
  }
};

// One block per (padded) query row; the row is read once (float4, coalesced) and kept in registers.
// Invalid rows (pad slots) are skipped: they only feed accumulator rows that the PV epilogue never
// stores.  Pad *columns* get P = 0.  NV4 = float4 per thread (row length <= 1024*NV4).
template <int NV4>
__global__ void __launch_bounds__(256) softmax_rows_kernel(ScoreParams p) {
  // blockIdx.x enumerates the rows of all un-fused problems back to back (softmax_row_begin prefix sums)
  int e = 0;
  while (e + 1 < p.nheads && (int)blockIdx.x >= p.softmax_row_begin[e + 1]) ++e;
  const AttnHead& h = p.h[e];
  const int qp = (int)blockIdx.x - p.softmax_row_begin[e];
  if (h.fused || qp >= h.ntt * 128) return;
  const int toh = qp / h.owp, owi = qp - toh * h.owp;
  if (owi >= h.ow || toh >= h.toh_total) return;
  const float* s = h.S + (size_t)qp * h.ldS;
  __half* pr = h.P + (size_t)qp * h.ldP;
  const int ncols = h.nk64 * 64;               // multiple of 64
  const int nvalid = h.toh_total * h.owp;
  const int owm = h.owp - 1, ow = h.ow;
  __shared__ float red[8];
  float4 v[NV4];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (c < ncols) {
      float4 a = *reinterpret_cast<const float4*>(s + c);
      for (int sp = 1; sp < h.splits; ++sp) {  // split-K slabs, fixed order -> deterministic
        const float4 b = *reinterpret_cast<const float4*>(s + (size_t)sp * h.slabS + c);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      v[i].x = (c + 0 < nvalid && ((c + 0) & owm) < ow) ? a.x : -INFINITY;
      v[i].y = (c + 1 < nvalid && ((c + 1) & owm) < ow) ? a.y : -INFINITY;
      v[i].z = (c + 2 < nvalid && ((c + 2) & owm) < ow) ? a.z : -INFINITY;
      v[i].w = (c + 3 < nvalid && ((c + 3) & owm) < ow) ? a.w : -INFINITY;
      m = fmaxf(m, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  const float sc = h.scale_log2e;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int c = (i * 256 + threadIdx.x) * 4;
    if (c < ncols) {
      // exp2f(-inf) = 0 for the masked slots
      const __half2 h01 = __floats2half2_rn(fast_exp2((v[i].x - m) * sc), fast_exp2((v[i].y - m) * sc));
      const __half2 h23 = __floats2half2_rn(fast_exp2((v[i].z - m) * sc), fast_exp2((v[i].w - m) * sc));
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      sum += (f01.x + f01.y) + (f23.x + f23.y);  // normalise by what the PV GEMM will actually multiply with
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h01);
      pk.y = *reinterpret_cast<const uint32_t*>(&h23);
      *reinterpret_cast<uint2*>(pr + c) = pk;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    h.rowsum[qp] = tot;
  }
}

// Direct heads: the row shift s_self[qp] = q_qp . k_qp over the token's ph*pw patch pixels x 64 channels (un-scaled, fp32), 0 for pad slots.
// One warp per padded token; a lane reads 16 bytes (8 channels), 8 lanes cover the 64 channels of a pixel, the warp 4 patch pixels per step,
// 4 steps unrolled (up to 16 pixels x 2 tensors x 128 B in flight per warp): reads straight from the NHWC projection buffer (the same
// addresses the TMA views of the score kernel read).  HBM/L2-bound: 2 x 128 B per pixel and head.
struct DiagParams {
  const __half* q[ATTN_MAX_HEADS];   // base of this problem's frames + q channel offset of the head
  const __half* k[ATTN_MAX_HEADS];
  AttnHead h[ATTN_MAX_HEADS];
  int row_begin[ATTN_MAX_HEADS + 1]; // prefix sums of ntt*128 over the direct problems (others: empty)
  int nheads, pitch, W;              // qkv pixel pitch (elements), feature-map width
};
__device__ __forceinline__ float dot8_f16(const uint4& a, const uint4& b) {
  const __half2* x = reinterpret_cast<const __half2*>(&a);
  const __half2* y = reinterpret_cast<const __half2*>(&b);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 u = __half22float2(x[i]), v = __half22float2(y[i]);
    acc = fmaf(u.x, v.x, fmaf(u.y, v.y, acc));
  }
  return acc;
}
__global__ void __launch_bounds__(256) attn_diag_kernel(DiagParams p) {
  const int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  int e = 0;
  while (e + 1 < p.nheads && gw >= p.row_begin[e + 1]) ++e;
  const AttnHead& h = p.h[e];
  const int qp = gw - p.row_begin[e];
  if (h.fused != 2 || qp >= h.ntt * 128) return;
  const int toh = qp / h.owp, owi = qp - toh * h.owp;
  float acc = 0.f;
  if (owi < h.ow && toh < h.toh_total) {
    const int tt = toh / h.oh, ohi = toh - tt * h.oh;
    const size_t pix0 = ((size_t)tt * h.oh * h.ph + (size_t)ohi * h.ph) * p.W + (size_t)owi * h.pw;   // H = oh * ph
    const int sub = lane >> 3, ch = (lane & 7) * 8, npos = h.npos;
    const __half* qb = p.q[e] + ch;
    const __half* kb = p.k[e] + ch;
    for (int i0 = 0; i0 < npos; i0 += 16) {
      uint4 a[4], b[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + 4 * u + sub;
        ok[u] = i < npos;
        const int py = ok[u] ? i / h.pw : 0, px = ok[u] ? i - py * h.pw : 0;
        const size_t o = (pix0 + (size_t)py * p.W + px) * p.pitch;
        a[u] = ok[u] ? *reinterpret_cast<const uint4*>(qb + o) : make_uint4(0, 0, 0, 0);
        b[u] = ok[u] ? *reinterpret_cast<const uint4*>(kb + o) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += dot8_f16(a[u], b[u]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if (lane == 0) h.rowshift[qp] = acc;
}

struct PVParams {
  CUtensorMap pmap[ATTN_MAX_HEADS], vmap[ATTN_MAX_HEADS];
  AttnHead h[ATTN_MAX_HEADS];
  int nheads, total_work, total_work2;
  int T, H, W;       // feature map geometry
  __half* out;       // NHWC fp16 [T,H,W,out_pitch]; entry s writes channels [coff[s], coff[s]+64) of the frames
  int out_pitch;     // starting at element offset out_off[s] (its window's first frame)
  int coff[ATTN_MAX_HEADS];
  long long out_off[ATTN_MAX_HEADS];
  const int* order;  // LPT schedule, see ScoreParams
  int order_stride;
  int* overflow;     // direct heads: range flag shared with the score kernel
};

struct PVPolicy {
  static constexpr int BN = 256;
  static constexpr int STAGES = 4;
  static constexpr int B_MN_MAJOR = 1;
  static constexpr int PROF_ID = 3;
  static constexpr bool EPI_SCRATCH = false;
  using Params = PVParams;
  struct Tile {
    int num_k, n_cols;
    int head, mi, ni, nvalid;
  };
  struct RowCtx {
    bool valid;
    float inv;
    __half* base;
  };
  __device__ static void prefetch(const Params& p) {
    for (int i = 0; i < p.nheads; ++i) {
      tma_prefetch_desc(&p.pmap[i]);
      tma_prefetch_desc(&p.vmap[i]);
    }
  }
  __device__ static int num_tiles(const Params& p) { return p.total_work; }
  __device__ static int tile_at(const Params& p, int who, int n_who, int i, int ntiles) {
    return (p.order) ? tc_tile_listed(p.order, p.order_stride, who, i) : tc_tile_round_robin(who, n_who, i, ntiles);
  }
  __device__ static Tile get_tile(const Params& p, int idx) {
    int hd = 0;
    while (hd + 1 < p.nheads && idx >= p.h[hd + 1].pv_work_begin) ++hd;
    const AttnHead& h = p.h[hd];
    idx -= h.pv_work_begin;
    Tile t;
    t.head = hd;
    t.ni = idx % h.pv_ntiles;
    t.mi = idx / h.pv_ntiles;
    t.nvalid = min(4, h.npos - t.ni * 4);
    t.num_k = h.nk64;
    t.n_cols = t.nvalid * 64;
    return t;
  }
  __device__ static void prefetch_tile(const Params&, const Tile&) {}
  struct LoadCtx {
    int px[4], py[4];  // patch pixel of each value position of this tile (fixed over k)
    int prow, rows;    // P row coordinate, value token rows per 64-token chunk
    uint32_t tx_bytes;
  };
  __device__ static LoadCtx load_begin(const Params& p, const Tile& t) {
    const AttnHead& h = p.h[t.head];
    LoadCtx lc;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pos = t.ni * 4 + j;
      lc.py[j] = pos / h.pw;
      lc.px[j] = pos - lc.py[j] * h.pw;
    }
    lc.prow = t.mi * 128;
    lc.rows = 64 / h.owp;
    lc.tx_bytes = (uint32_t)(TC_A_BYTES + t.nvalid * 8192);
    return lc;
  }
  __device__ static void load(const Params& p, const Tile& t, LoadCtx& lc, int k, uint32_t sA, uint32_t sB, uint32_t bar) {
    mbar_expect_tx(bar, lc.tx_bytes);
    tma_load_2d(sA, &p.pmap[t.head], bar, k * 64, lc.prow);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < t.nvalid) tma_load_5d(sB + j * 8192, &p.vmap[t.head], bar, 0, lc.px[j], 0, lc.py[j], k * lc.rows);
  }
  __device__ static RowCtx row_begin(const Params& p, const Tile& t, int row) {
    const AttnHead& h = p.h[t.head];
    RowCtx c;
    const int qp = t.mi * 128 + row;
    const int toh = qp / h.owp, owi = qp - toh * h.owp;
    c.valid = owi < h.ow && toh < h.toh_total;
    c.inv = 0.f;
    c.base = nullptr;
    if (c.valid) {
      float rs;
      if (h.fused) {
        const float* ps = h.rowsum_part + (size_t)qp * h.npairs;
        // direct heads: a row whose largest P is below 2^-12 sits in fp16's subnormals — ask for the exact path
        // (checked after the sum below)
        rs = ps[0];
        for (int j = 1; j < h.npairs; ++j) rs += ps[j];  // fixed order: deterministic
      } else {
        rs = h.rowsum[qp];
      }
      if (h.fused == 2 && !(rs >= 0.5f) && p.overflow) *p.overflow = 2;   // a direct row always holds its own 1: anything else is a defect
      c.inv = 1.0f / rs;
      const int tt = toh / h.oh, ohi = toh - tt * h.oh;
      c.base = p.out + p.out_off[t.head] + (((size_t)tt * p.H + ohi * h.ph) * p.W + owi * h.pw) * p.out_pitch + p.coff[t.head];
    }
    return c;
  }
  __device__ static void row_end(const Params&, const Tile&, RowCtx&, int) {}
  __device__ static void epilogue(const Params& p, const Tile& t, RowCtx& c, int row, int col0, float* v, float*) {
    if (!c.valid) return;
    const AttnHead& h = p.h[t.head];
    const int pos = t.ni * 4 + (col0 >> 6);
    const int py = pos / h.pw, px = pos - py * h.pw;
    __half* o = c.base + ((size_t)py * p.W + px) * p.out_pitch + (col0 & 63);
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      __align__(16) __half2 hh[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) hh[j] = __floats2half2_rn(v[i + 2 * j] * c.inv, v[i + 2 * j + 1] * c.inv);
      *reinterpret_cast<uint4*>(o + i) = *reinterpret_cast<const uint4*>(hh);
    }
  }
};

// ---- CTA-pair (cta_group::2) versions: 256 queries x 256 keys / 256 queries x 4 positions per pair -------
// Same math and epilogues as ScorePolicy / PVPolicy; each CTA loads its own 128 query rows and its own half
// of the B operand (one 128-key tile / two of the four value positions): 32 KB instead of 48 KB per stage
// and six stages in flight instead of four.
struct Score2Policy {
  static constexpr int STAGES = 6;
  static constexpr int B_MN_MAJOR = 0;
  static constexpr int PROF_ID = 5;
  static constexpr bool EPI_SCRATCH = true;
  using Params = ScoreParams;
  using Tile = ScorePolicy::Tile;
  using RowCtx = ScorePolicy::RowCtx;
  __device__ static void prefetch(const Params& p) { ScorePolicy::prefetch(p); }
  __device__ static int num_tiles(const Params& p) { return p.pass ? p.totalB2 : p.total_work2; }
  __device__ static int tile_at(const Params& p, int who, int n_who, int i, int ntiles) {
    return (p.order && p.pass == 0) ? tc_tile_listed(p.order, p.order_stride, who, i) : tc_tile_round_robin(who, n_who, i, ntiles);
  }
  __device__ static Tile get_tile(const Params& p, int idx, uint32_t rank) {
    int hd = 0;
    if (p.pass) {
      while (hd + 1 < p.nheads && idx >= p.h[hd + 1].scoreB_begin2) ++hd;
      idx -= p.h[hd].scoreB_begin2;
    } else {
      while (hd + 1 < p.nheads && idx >= p.h[hd + 1].score_work_begin2) ++hd;
      idx -= p.h[hd].score_work_begin2;
    }
    const AttnHead& h = p.h[hd];
    Tile t;
    t.head = hd;
    const int sp = idx % h.splits;
    idx /= h.splits;
    const int nkt2 = (h.ntt + 1) >> 1;
    t.kj = idx % nkt2;
    t.qi = (idx / nkt2) * 2 + (int)rank;   // this CTA's 128-query tile (may be == ntt: dummy, zero-filled)
    t.kbeg = sp * h.chunks_per_split;
    t.num_k = min(h.chunks_per_split, h.npos - t.kbeg);
    t.nkt = min(2, h.ntt - 2 * t.kj);
    t.n_cols = (t.qi < h.ntt) ? 128 * t.nkt : 0;
    t.split = sp;
    return t;
  }
  __device__ static void prefetch_tile(const Params&, const Tile&) {}
  using LoadCtx = ScorePolicy::LoadCtx;
  __device__ static LoadCtx load_begin(const Params& p, const Tile& t, uint32_t rank) {
    LoadCtx lc = ScorePolicy::load_begin(p, t);
    lc.krow0 = (2 * t.kj + (int)rank) * (128 / p.h[t.head].owp);  // this CTA's key tile
    lc.tx_bytes = (uint32_t)((2 + t.nkt) * TC_A_BYTES);          // registered by the leader for both CTAs
    return lc;
  }
  __device__ static void load(const Params& p, const Tile& t, LoadCtx& lc, int, uint32_t sA, uint32_t sB, uint32_t local_full,
                              uint32_t leader_full, uint32_t rank) {
    if (rank == 0) mbar_expect_tx(local_full, lc.tx_bytes);
    tma_load_5d_2sm(sA, &p.qmap[t.head], leader_full, 0, lc.px, 0, lc.py, lc.qrow);
    if ((int)rank < t.nkt) tma_load_5d_2sm(sB, &p.kmap[t.head], leader_full, 0, lc.px, 0, lc.py, lc.krow0);
    if (++lc.px == lc.pw) { lc.px = 0; ++lc.py; }
  }
  __device__ static RowCtx row_begin(const Params& p, const Tile& t, int row) { return ScorePolicy::row_begin(p, t, row); }
  __device__ static void epilogue(const Params& p, const Tile& t, RowCtx& c, int row, int col0, float* v, float* scr) {
    ScorePolicy::epilogue(p, t, c, row, col0, v, scr);
  }
  __device__ static void row_end(const Params& p, const Tile& t, RowCtx& c, int row) { ScorePolicy::row_end(p, t, c, row); }
};

struct PV2Policy {
  static constexpr int STAGES = 6;
  static constexpr int B_MN_MAJOR = 1;
  static constexpr int PROF_ID = 6;
  static constexpr bool EPI_SCRATCH = false;
  using Params = PVParams;
  using Tile = PVPolicy::Tile;
  using RowCtx = PVPolicy::RowCtx;
  __device__ static void prefetch(const Params& p) { PVPolicy::prefetch(p); }
  __device__ static int num_tiles(const Params& p) { return p.total_work2; }
  __device__ static int tile_at(const Params& p, int who, int n_who, int i, int ntiles) {
    return (p.order) ? tc_tile_listed(p.order, p.order_stride, who, i) : tc_tile_round_robin(who, n_who, i, ntiles);
  }
  __device__ static Tile get_tile(const Params& p, int idx, uint32_t rank) {
    int hd = 0;
    while (hd + 1 < p.nheads && idx >= p.h[hd + 1].pv_work_begin2) ++hd;
    const AttnHead& h = p.h[hd];
    idx -= h.pv_work_begin2;
    Tile t;
    t.head = hd;
    t.ni = idx % h.pv_ntiles;
    t.mi = (idx / h.pv_ntiles) * 2 + (int)rank;  // dummy when == ntt: P rows beyond the tensor are zero-filled
    t.nvalid = min(4, h.npos - t.ni * 4);
    t.num_k = h.nk64;
    t.n_cols = t.nvalid * 64;
    return t;
  }
  __device__ static void prefetch_tile(const Params&, const Tile&) {}
  using LoadCtx = PVPolicy::LoadCtx;
  __device__ static LoadCtx load_begin(const Params& p, const Tile& t, uint32_t rank) {
    LoadCtx lc = PVPolicy::load_begin(p, t);
    if (rank) {  // this CTA owns value positions 2 and 3 of the tile: keep them in slots 0 and 1
      lc.px[0] = lc.px[2]; lc.py[0] = lc.py[2];
      lc.px[1] = lc.px[3]; lc.py[1] = lc.py[3];
    }
    lc.tx_bytes = (uint32_t)(2 * TC_A_BYTES + t.nvalid * 8192);  // registered by the leader for both CTAs
    return lc;
  }
  __device__ static void load(const Params& p, const Tile& t, LoadCtx& lc, int k, uint32_t sA, uint32_t sB, uint32_t local_full,
                              uint32_t leader_full, uint32_t rank) {
    if (rank == 0) mbar_expect_tx(local_full, lc.tx_bytes);
    tma_load_2d_2sm(sA, &p.pmap[t.head], leader_full, k * 64, lc.prow);
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (2 * (int)rank + j < t.nvalid)
        tma_load_5d_2sm(sB + j * 8192, &p.vmap[t.head], leader_full, 0, lc.px[j], 0, lc.py[j], k * lc.rows);
  }
  __device__ static RowCtx row_begin(const Params& p, const Tile& t, int row) { return PVPolicy::row_begin(p, t, row); }
  __device__ static void epilogue(const Params& p, const Tile& t, RowCtx& c, int row, int col0, float* v, float* scr) {
    PVPolicy::epilogue(p, t, c, row, col0, v, scr);
  }
  __device__ static void row_end(const Params&, const Tile&, RowCtx&, int) {}
};

}  // namespace vsr
