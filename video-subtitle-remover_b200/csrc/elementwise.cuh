// HBM-bound elementwise kernels either side of the network: strip pre-processing (cv2-exact u8
// bilinear down-scale), the 3->64 stride-2 stem conv, bilinear x2 (align_corners=True), and the
// post-processing (cv2-exact up-scale of the composites, u8 truncation, channel swap, mask select).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vsr {

// cv::resize INTER_LINEAR tap table for one axis (built on the host, see resize_tables() in engine.cu)
struct ResizeTaps {
  const int* i0;      // first source index
  const int* i1;      // second source index
  const float* a;     // fractional weight of i1 (float path)
  const short* w0;    // rint((1-a)*2048)  (u8 path)
  const short* w1;    // rint(a*2048)
};

// A3/A4: crop is implicit (src points at the strip rows), cv2.resize u8 -> [T, dh, dw] RGBA8 with
// channels swapped to RGB (Stack(), utils/sttn_utils.py:73).  One thread per output pixel.
__global__ void __launch_bounds__(256) strip_downscale_kernel(const uint8_t* __restrict__ src, size_t src_frame_stride, int sw,
                                                              int sh, uint8_t* __restrict__ dst, int dw, int dh, int T,
                                                              ResizeTaps tx, ResizeTaps ty) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int t = blockIdx.z;
  if (x >= dw) return;
  const uint8_t* f = src + (size_t)t * src_frame_stride;
  const int x0 = tx.i0[x] * 3, x1 = tx.i1[x] * 3;
  const int wx0 = tx.w0[x], wx1 = tx.w1[x];
  const int wy0 = ty.w0[y], wy1 = ty.w1[y];
  const uint8_t* r0 = f + (size_t)ty.i0[y] * sw * 3;
  const uint8_t* r1 = f + (size_t)ty.i1[y] * sw * 3;
  uchar4 o;
  uint8_t* op = reinterpret_cast<uint8_t*>(&o);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int s0 = r0[x0 + c] * wx0 + r0[x1 + c] * wx1;
    const int s1 = r1[x0 + c] * wx0 + r1[x1 + c] * wx1;
    const int v = (((wy0 * (s0 >> 4)) >> 16) + ((wy1 * (s1 >> 4)) >> 16) + 2) >> 2;
    op[2 - c] = (uint8_t)min(max(v, 0), 255);  // BGR -> RGB
  }
  op[3] = 0;
  reinterpret_cast<uchar4*>(dst)[((size_t)t * dh + y) * dw + x] = o;
}

// cv2.resize of the single-channel mask strip (sttn_det_inpaint.py:73), same fixed-point arithmetic.
__global__ void __launch_bounds__(256) mask_downscale_kernel(const uint8_t* __restrict__ src, int sw, int sh, uint8_t* __restrict__ dst,
                                                             int dw, int dh, ResizeTaps tx, ResizeTaps ty) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= dw) return;
  const uint8_t* r0 = src + (size_t)ty.i0[y] * sw;
  const uint8_t* r1 = src + (size_t)ty.i1[y] * sw;
  const int s0 = r0[tx.i0[x]] * tx.w0[x] + r0[tx.i1[x]] * tx.w1[x];
  const int s1 = r1[tx.i0[x]] * tx.w0[x] + r1[tx.i1[x]] * tx.w1[x];
  const int v = (((ty.w0[y] * (s0 >> 4)) >> 16) + ((ty.w1[y] * (s1 >> 4)) >> 16) + 2) >> 2;
  dst[(size_t)y * dw + x] = (uint8_t)min(max(v, 0), 255);
}

// A5 first layer (auto_sttn.py:76-77): conv3x3 stride 2 pad 1, 3 -> 64, LeakyReLU(0.2), computed in
// fp32 from the RGBA8 image with the `/255*2-1` normalisation of sttn_auto_inpaint.py:128 folded in.
// in [T,H,W] RGBA8 -> out NHWC fp16 [T,H/2,W/2,64].  Block = 64 output pixels x 4 channel groups.
// sttn-det: `det_mask` [H,W] u8 (the resized mask); pixels with mask/255 > 0.5 enter the encoder as 0
// (feats*(1-masks_tensor), sttn_det_inpaint.py:134,143).
__global__ void __launch_bounds__(256) stem_conv_kernel(const uchar4* __restrict__ in, int H, int W, const float* __restrict__ wgt,
                                                        const float* __restrict__ bias, __half* __restrict__ out, int total_pix,
                                                        const uint8_t* __restrict__ det_mask) {
  __shared__ float sw[27 * 64];  // [tap*3+ci][co]
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 27 * 64; i += 256) sw[i] = wgt[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int OW = W >> 1, OH = H >> 1;
  const int pix = blockIdx.x * 64 + (threadIdx.x & 63);
  const int cg = threadIdx.x >> 6;  // 16 output channels each
  if (pix >= total_pix) return;
  const int ox = pix % OW;
  const int oy = (pix / OW) % OH;
  const int t = pix / (OW * OH);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = sb[cg * 16 + i];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - 1;
      if (ix < 0 || ix >= W) continue;
      const uchar4 px = in[((size_t)t * H + iy) * W + ix];
      float v[3] = {(float)px.x / 255.0f * 2.0f - 1.0f, (float)px.y / 255.0f * 2.0f - 1.0f,
                    (float)px.z / 255.0f * 2.0f - 1.0f};
      if (det_mask && det_mask[(size_t)iy * W + ix] >= 128) v[0] = v[1] = v[2] = 0.f;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float* wr = sw + ((ky * 3 + kx) * 3 + ci) * 64 + cg * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(v[ci], wr[i], acc[i]);
      }
    }
  }
  __align__(16) __half2 h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = acc[2 * i], b = acc[2 * i + 1];
    a = a > 0.f ? a : 0.2f * a;
    b = b > 0.f ? b : 0.2f * b;
    h[i] = __floats2half2_rn(a, b);
  }
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)pix * 64 + cg * 16);
  o[0] = reinterpret_cast<const uint4*>(h)[0];
  o[1] = reinterpret_cast<const uint4*>(h)[1];
}

// deconv's F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) (auto_sttn.py:124-126)
// on NHWC fp16; one thread = 8 channels of one output pixel.
__global__ void __launch_bounds__(256) upsample2x_kernel(const __half* __restrict__ in, int T, int h, int w, int C,
                                                         __half* __restrict__ out) {
  const int cg = C >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * (2 * h) * (2 * w) * cg;
  if (idx >= total) return;
  const int c8 = idx % cg;
  size_t r = idx / cg;
  const int ox = r % (2 * w);
  r /= (2 * w);
  const int oy = r % (2 * h);
  const int t = r / (2 * h);
  const float sy = (float)(h - 1) / (float)(2 * h - 1), sx = (float)(w - 1) / (float)(2 * w - 1);
  const float fy = sy * oy, fx = sx * ox;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const __half* base = in + (size_t)t * h * w * C + c8 * 8;
  const uint4 a = *reinterpret_cast<const uint4*>(base + ((size_t)y0 * w + x0) * C);
  const uint4 b = *reinterpret_cast<const uint4*>(base + ((size_t)y0 * w + x1) * C);
  const uint4 c = *reinterpret_cast<const uint4*>(base + ((size_t)y1 * w + x0) * C);
  const uint4 d = *reinterpret_cast<const uint4*>(base + ((size_t)y1 * w + x1) * C);
  const __half2* pa = reinterpret_cast<const __half2*>(&a);
  const __half2* pb = reinterpret_cast<const __half2*>(&b);
  const __half2* pc = reinterpret_cast<const __half2*>(&c);
  const __half2* pd = reinterpret_cast<const __half2*>(&d);
  __align__(16) __half2 o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __half22float2(pa[i]), fb = __half22float2(pb[i]), fc = __half22float2(pc[i]), fd = __half22float2(pd[i]);
    const float u = hy * (hx * fa.x + lx * fb.x) + ly * (hx * fc.x + lx * fd.x);
    const float v = hy * (hx * fa.y + lx * fb.y) + ly * (hx * fc.y + lx * fd.y);
    o[i] = __floats2half2_rn(u, v);
  }
  *reinterpret_cast<uint4*>(out + (((size_t)t * 2 * h + oy) * 2 * w + ox) * C + c8 * 8) = *reinterpret_cast<const uint4*>(o);
}

// Scene-cut detection (backend/scenedetect ContentDetector through subtitle_detect.py:158-170): one pass over a decoded BGR
// frame that (1) down-scales it like the reference's decode thread (cv2.resize INTER_LINEAR in its 11-bit fixed point, or the rounded 2x2
// mean OpenCV substitutes for an exact 2x down-scale; `mode` 0 = none, 1 = linear, 2 = 2x2 mean), (2) converts the pixel to 8-bit HSV with
// OpenCV's 12-bit tables (RGB2HSV_b, hue range 180), (3) stores it for the next frame and (4) accumulates |delta| of hue, saturation and
// value against the previous frame's pixel into three 64-bit integer sums — integers, so the frame score is bit-exact on the host.
// HBM-bound: the frame is read once (6.2 MB at 1080p), the small HSV images stay in L2.
__global__ void __launch_bounds__(256) scene_hsv_diff_kernel(const uint8_t* __restrict__ bgr, int sw, int sh, int dw, int dh, int mode,
                                                             ResizeTaps tx, ResizeTaps ty, const int* __restrict__ sdiv,
                                                             const int* __restrict__ hdiv, const uchar4* __restrict__ prev,
                                                             uchar4* __restrict__ cur, int have_prev, unsigned long long* __restrict__ sums) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  int dh_ = 0, ds_ = 0, dv_ = 0;
  if (x < dw) {
    int c3[3];
    if (mode == 0) {
      const uint8_t* p = bgr + ((size_t)y * sw + x) * 3;
      c3[0] = p[0]; c3[1] = p[1]; c3[2] = p[2];
    } else if (mode == 2) {
      const uint8_t* p0 = bgr + ((size_t)(2 * y) * sw + 2 * x) * 3;
      const uint8_t* p1 = p0 + (size_t)sw * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) c3[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
    } else {
      const int x0 = tx.i0[x] * 3, x1 = tx.i1[x] * 3;
      const int wx0 = tx.w0[x], wx1 = tx.w1[x], wy0 = ty.w0[y], wy1 = ty.w1[y];
      const uint8_t* r0 = bgr + (size_t)ty.i0[y] * sw * 3;
      const uint8_t* r1 = bgr + (size_t)ty.i1[y] * sw * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int s0 = r0[x0 + c] * wx0 + r0[x1 + c] * wx1;
        const int s1 = r1[x0 + c] * wx0 + r1[x1 + c] * wx1;
        c3[c] = min(max((((wy0 * (s0 >> 4)) >> 16) + ((wy1 * (s1 >> 4)) >> 16) + 2) >> 2, 0), 255);
      }
    }
    const int b = c3[0], g = c3[1], r = c3[2];
    const int v = max(max(b, g), r), vmin = min(min(b, g), r), diff = v - vmin;
    const int s = (diff * sdiv[v] + (1 << 11)) >> 12;
    int h = (v == r) ? (g - b) : (v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff);
    h = (h * hdiv[diff] + (1 << 11)) >> 12;
    h += h < 0 ? 180 : 0;
    const size_t o = (size_t)y * dw + x;
    cur[o] = make_uchar4((unsigned char)h, (unsigned char)s, (unsigned char)v, 0);
    if (have_prev) {
      const uchar4 q = prev[o];
      dh_ = abs(h - (int)q.x); ds_ = abs(s - (int)q.y); dv_ = abs(v - (int)q.z);
    }
  }
  if (!have_prev) return;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    dh_ += __shfl_xor_sync(0xffffffffu, dh_, o);
    ds_ += __shfl_xor_sync(0xffffffffu, ds_, o);
    dv_ += __shfl_xor_sync(0xffffffffu, dv_, o);
  }
  __shared__ int red[3][8];
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dh_; red[1][threadIdx.x >> 5] = ds_; red[2][threadIdx.x >> 5] = dv_; }
  __syncthreads();
  if (threadIdx.x < 3) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
    atomicAdd(&sums[threadIdx.x], (unsigned long long)t);
  }
}

// Sharded windows (vsr_sttn_shard_*): every window was decoded into its own slot `preds[slot][local frame]` (quantised, unblended);
// this replays the running 0.5 / 0.5 blend of sttn_auto_inpaint.py:159-162 per frame in schedule order — the same two multiplies
// and one add per visit as the CONV_FINAL epilogue, so the result is bit-identical to the single-GPU chunk.
// visit_tab[f * 4] = number of visits (<= 3), visit_tab[f * 4 + 1 + k] = slot * 32 + local frame index of visit k.
__global__ void __launch_bounds__(256) blend_preds_kernel(float* __restrict__ comps, const float* __restrict__ preds,
                                                          const int* __restrict__ visit_tab, size_t frame_elems) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (e >= frame_elems) return;
  const int n = visit_tab[f * 4];
  if (n <= 0) return;
  float c = preds[(size_t)visit_tab[f * 4 + 1] * frame_elems + e];
  for (int k = 1; k < n; ++k) c = c * 0.5f + preds[(size_t)visit_tab[f * 4 + 1 + k] * frame_elems + e] * 0.5f;
  comps[(size_t)f * frame_elems + e] = c;
}

// A3 tail (sttn_auto_inpaint.py:86-91 / 312-315): comp = cv2.resize(comp, (W, split_h)) [u8 fixed-point
// path when the frame was decoded once, float path otherwise] -> astype(uint8) -> RGB->BGR ->
// strip = mask ? comp : strip.  comps [T, ch, cw, 3] fp32 RGB; strips [T, sh, sw, 3] u8 BGR in place.
__global__ void __launch_bounds__(256) strip_composite_kernel(const float* __restrict__ comps, int cw, int ch,
                                                              const int* __restrict__ visits, const uint8_t* __restrict__ mask,
                                                              int mask_pitch, uint8_t* __restrict__ strips, size_t strip_frame_stride,
                                                              int sw, int sh, int T, ResizeTaps tx, ResizeTaps ty) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int t = blockIdx.z;
  if (x >= sw) return;
  if (mask && mask[(size_t)y * mask_pitch + x] == 0) return;  // mask == nullptr (sttn-det): the whole strip is replaced
  const float* f = comps + (size_t)t * ch * cw * 3;
  const int x0 = tx.i0[x], x1 = tx.i1[x], y0 = ty.i0[y], y1 = ty.i1[y];
  const float* p00 = f + ((size_t)y0 * cw + x0) * 3;
  const float* p01 = f + ((size_t)y0 * cw + x1) * 3;
  const float* p10 = f + ((size_t)y1 * cw + x0) * 3;
  const float* p11 = f + ((size_t)y1 * cw + x1) * 3;
  uint8_t* o = strips + (size_t)t * strip_frame_stride + ((size_t)y * sw + x) * 3;
  if (visits[t] <= 1) {
    const int wx0 = tx.w0[x], wx1 = tx.w1[x], wy0 = ty.w0[y], wy1 = ty.w1[y];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int s0 = (int)p00[c] * wx0 + (int)p01[c] * wx1;
      const int s1 = (int)p10[c] * wx0 + (int)p11[c] * wx1;
      const int v = (((wy0 * (s0 >> 4)) >> 16) + ((wy1 * (s1 >> 4)) >> 16) + 2) >> 2;
      o[2 - c] = (uint8_t)min(max(v, 0), 255);
    }
  } else {
    const float ax = tx.a[x], ay = ty.a[y];
    const float bx = 1.0f - ax, by = 1.0f - ay;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // cv::resize float path: horizontal then vertical, separate roundings (no contraction)
      const float h0 = __fadd_rn(__fmul_rn(p00[c], bx), __fmul_rn(p01[c], ax));
      const float h1 = __fadd_rn(__fmul_rn(p10[c], bx), __fmul_rn(p11[c], ax));
      const float v = __fadd_rn(__fmul_rn(h0, by), __fmul_rn(h1, ay));
      o[2 - c] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);  // astype(uint8): truncation
    }
  }
}

}  // namespace vsr
