// Generic NHWC fp16 operators for graph-driven networks (the DBNet text detector): the
// dense convolutions go through the tcgen05 implicit-GEMM kernel of conv_igemm.cuh; this file holds the
// HBM-bound rest (depthwise conv, small-Cin stem, space-to-depth, element-wise, nearest up-sampling, max
// pooling, channel copies for concat, 2x2 stride-2 transposed conv, image normalisation).
// Every tensor is [T, H, W, Cp] with the channel pitch Cp a multiple of 8 (64 for tensors that feed a TMA
// conv); channels >= C are padding that the consumers multiply with zero weights.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vsr {

enum RtEltOp : int { RT_ADD = 0, RT_RELU = 1, RT_ADD_RELU = 2, RT_SIGMOID = 3, RT_AFFINE = 4, RT_AFFINE_RELU = 5, RT_SCALE = 6,
                     RT_AVG2 = 7 };

// Per-tensor power-of-two scaling: a tensor stores value * s so that networks without normalisation
// (the detector's LKPAN neck reaches |x| ~ 1e5) stay inside fp16.  A layer computes out = acc * alpha + bias *
// bias_scale with alpha = s_out / s_in, bias_scale = s_out, and raises *overflow when a stored value leaves fp16.
struct RtScale {
  float alpha, bias_scale;
  int* overflow;
};

__device__ __forceinline__ void rt_store8(const float* acc, const float* __restrict__ bias, const RtScale& sc, int relu, __half* dst) {
  __align__(16) __half2 o[4];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = fmaf(acc[2 * j], sc.alpha, bias[2 * j] * sc.bias_scale);
    float b = fmaf(acc[2 * j + 1], sc.alpha, bias[2 * j + 1] * sc.bias_scale);
    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
    bad |= !(fabsf(a) <= 65504.f) | !(fabsf(b) <= 65504.f);
    o[j] = __floats2half2_rn(a, b);
  }
  if (bad && sc.overflow) *sc.overflow = 1;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
}

// 8 channels per thread.  a, b, out may alias.  ADD / ADD_RELU: out = a*alpha + b*beta (operands at different scales);
// SIGMOID: out = sigmoid(a*alpha); AFFINE: out = a*scale[c] + shift[c]; SCALE: out = a*alpha + beta; AVG2: out = (a + b) * alpha.
__global__ void __launch_bounds__(256) rt_elementwise_kernel(int op, const __half* __restrict__ a, const __half* __restrict__ b,
                                                             __half* __restrict__ out, size_t n8, int cp, const float* __restrict__ scale,
                                                             const float* __restrict__ shift, float alpha, float beta, int* overflow) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 va = reinterpret_cast<const uint4*>(a)[i];
  uint4 vb = make_uint4(0, 0, 0, 0);
  if (op == RT_ADD || op == RT_ADD_RELU || op == RT_AVG2) vb = reinterpret_cast<const uint4*>(b)[i];
  const __half2* pa = reinterpret_cast<const __half2*>(&va);
  const __half2* pb = reinterpret_cast<const __half2*>(&vb);
  const int c0 = (int)((i * 8) % (size_t)cp);
  __align__(16) __half2 o[4];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 x = __half22float2(pa[j]);
    const float2 y = __half22float2(pb[j]);
    switch (op) {
      case RT_ADD: x.x = fmaf(x.x, alpha, y.x * beta); x.y = fmaf(x.y, alpha, y.y * beta); break;
      case RT_RELU: x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); break;
      case RT_ADD_RELU: x.x = fmaxf(fmaf(x.x, alpha, y.x * beta), 0.f); x.y = fmaxf(fmaf(x.y, alpha, y.y * beta), 0.f); break;
      case RT_SIGMOID: x.x = 1.f / (1.f + __expf(-x.x * alpha)); x.y = 1.f / (1.f + __expf(-x.y * alpha)); break;
      case RT_AFFINE:
      case RT_AFFINE_RELU:
        x.x = x.x * scale[c0 + 2 * j] + shift[c0 + 2 * j];
        x.y = x.y * scale[c0 + 2 * j + 1] + shift[c0 + 2 * j + 1];
        if (op == RT_AFFINE_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); }
        break;
      case RT_SCALE: x.x = x.x * alpha + beta; x.y = x.y * alpha + beta; break;
      case RT_AVG2: x.x = (x.x + y.x) * alpha; x.y = (x.y + y.y) * alpha; break;
    }
    bad |= !(fabsf(x.x) <= 65504.f) | !(fabsf(x.y) <= 65504.f);
    o[j] = __floats2half2_rn(x.x, x.y);
  }
  if (bad && overflow) *overflow = 1;
  reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(o);
}

// ---- LAMA ---------------------------------------------------------------------------------------
// out[oy][ox] = in[r(oy - top)][r(ox - left)], r = reflection without repeating the edge (nn.ReflectionPad2d /
// padding_mode='reflect'); mode 0 writes zeros outside instead.  8 channels per thread.
__global__ void __launch_bounds__(256) rt_pad_kernel(const __half* __restrict__ in, int T, int H, int W, int cp, __half* __restrict__ out,
                                                     int OH, int OW, int top, int left, int reflect) {
  const int c8n = cp >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)T * OH * OW * c8n) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = r % OW;
  r /= OW;
  const int oy = r % OH;
  const int t = r / OH;
  int y = oy - top, x = ox - left;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (reflect) {
    y = y < 0 ? -y : y;
    x = x < 0 ? -x : x;
    y = y >= H ? 2 * (H - 1) - y : y;
    x = x >= W ? 2 * (W - 1) - x : x;
    y = min(max(y, 0), H - 1);   // alignment rows beyond one reflection are never read by a kept output
    x = min(max(x, 0), W - 1);
    v = *reinterpret_cast<const uint4*>(in + (((size_t)t * H + y) * W + x) * cp + c8 * 8);
  } else if (y >= 0 && y < H && x >= 0 && x < W) {
    v = *reinterpret_cast<const uint4*>(in + (((size_t)t * H + y) * W + x) * cp + c8 * 8);
  }
  *reinterpret_cast<uint4*>(out + idx * 8) = v;
}

// zero insertion for a stride-2 transposed conv: out[2y][2x] = in[y][x], 0 elsewhere ([T,2H,2W,cp])
__global__ void __launch_bounds__(256) rt_zero_upsample2x_kernel(const __half* __restrict__ in, int T, int H, int W, int cp,
                                                                 __half* __restrict__ out) {
  const int c8n = cp >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)T * 4 * H * W * c8n) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = r % (2 * W);
  r /= 2 * W;
  const int oy = r % (2 * H);
  const int t = r / (2 * H);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (((ox | oy) & 1) == 0) v = *reinterpret_cast<const uint4*>(in + (((size_t)t * H + (oy >> 1)) * W + (ox >> 1)) * cp + c8 * 8);
  *reinterpret_cast<uint4*>(out + idx * 8) = v;
}

// out[p][co + c] = f(a[p][ca + c] * alpha + b[p][cb + c] * beta), c < 8*c8n: add / add+relu between channel slices of
// tensors with different pitches (op 0 or 2 of RtEltOp)
__global__ void __launch_bounds__(256) rt_add_slices_kernel(int relu, const __half* __restrict__ a, int pa, const __half* __restrict__ b, int pb,
                                                            __half* __restrict__ out, int po, int c8n, size_t pixels, float alpha, float beta,
                                                            int* overflow) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pixels * c8n) return;
  const int c8 = idx % c8n;
  const size_t p = idx / c8n;
  const uint4 va = *reinterpret_cast<const uint4*>(a + p * pa + c8 * 8);
  const uint4 vb = *reinterpret_cast<const uint4*>(b + p * pb + c8 * 8);
  const __half2* ha = reinterpret_cast<const __half2*>(&va);
  const __half2* hb = reinterpret_cast<const __half2*>(&vb);
  __align__(16) __half2 o[4];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = __half22float2(ha[j]), y = __half22float2(hb[j]);
    float u = fmaf(x.x, alpha, y.x * beta), w = fmaf(x.y, alpha, y.y * beta);
    if (relu) { u = fmaxf(u, 0.f); w = fmaxf(w, 0.f); }
    bad |= !(fabsf(u) <= 65504.f) | !(fabsf(w) <= 65504.f);
    o[j] = __floats2half2_rn(u, w);
  }
  if (bad && overflow) *overflow = 1;
  *reinterpret_cast<uint4*>(out + p * po + c8 * 8) = *reinterpret_cast<const uint4*>(o);
}

// residual stream with an fp32 master copy: x32 (+)= y16, x16 = half(x32); init = 1 takes x16 as the initial value of x32.
// 18 residual blocks then round the stream to fp16 once per block for the next conv instead of accumulating in fp16.
__global__ void __launch_bounds__(256) rt_residual_add_kernel(float* __restrict__ x32, const __half* __restrict__ y16, __half* __restrict__ x16,
                                                              size_t n8, int init, int* overflow) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 vy = reinterpret_cast<const uint4*>(y16)[i];
  const __half2* hy = reinterpret_cast<const __half2*>(&vy);
  float v[8];
  if (init) {
    const uint4 vx = reinterpret_cast<const uint4*>(x16)[i];
    const __half2* hx = reinterpret_cast<const __half2*>(&vx);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(hx[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
  } else {
    const float4 a = reinterpret_cast<const float4*>(x32)[2 * i], b = reinterpret_cast<const float4*>(x32)[2 * i + 1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __align__(16) __half2 o[4];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(hy[j]);
    v[2 * j] += f.x;
    v[2 * j + 1] += f.y;
    bad |= !(fabsf(v[2 * j]) <= 65504.f) | !(fabsf(v[2 * j + 1]) <= 65504.f);
    o[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
  }
  if (bad && overflow) *overflow = 1;
  reinterpret_cast<float4*>(x32)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(x32)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
  reinterpret_cast<uint4*>(x16)[i] = *reinterpret_cast<const uint4*>(o);
}

// fp16 NHWC <-> fp32 planar staging around cuFFT (its half-precision transforms are power-of-two only, and its kernels
// want unit-stride signals: the first version handed it the interleaved-channel layout through istride = channels and spent
// 40 % of the LAMA frame in strided FFT kernels).  NHWC side: [T][P][Kc][E] halves (P pixels, Kc channel units of E = 1
// real or E = 2 (re, im) values); planar side: [(t*Kc + kc)][P][E] floats = one contiguous signal per channel.  32 x 32
// (pixel x unit) tiles go through shared memory so that both sides are accessed along their contiguous axis.
template <int E>
__global__ void __launch_bounds__(256) rt_nhwc_to_planar_kernel(const __half* __restrict__ in, float* __restrict__ out, size_t P, int Kc) {
  __shared__ float tile[32][32 * E + 1];
  const int t = blockIdx.z;
  const size_t p0 = (size_t)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  const int W = 32 * E;
  for (int i = threadIdx.x; i < 32 * W; i += 256) {       // read: consecutive threads walk the channel axis of one pixel
    const int pp = i / W, e = i - pp * W;
    const size_t p = p0 + pp;
    const int ku = k0 + e / E;
    tile[pp][e] = (p < P && ku < Kc) ? __half2float(in[((size_t)t * P + p) * Kc * E + (size_t)k0 * E + e]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * W; i += 256) {       // write: consecutive threads walk the pixel axis of one channel
    const int ku = i / W, r = i - ku * W;                   // r = pixel * E + e
    const size_t p = p0 + r / E;
    if (p < P && k0 + ku < Kc) out[(((size_t)t * Kc + k0 + ku) * P + p) * E + (r % E)] = tile[r / E][ku * E + (r % E)];
  }
}
template <int E>
__global__ void __launch_bounds__(256) rt_planar_to_nhwc_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t P, int Kc, float mul,
                                                                int* overflow) {
  __shared__ float tile[32][32 * E + 1];
  const int t = blockIdx.z;
  const size_t p0 = (size_t)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  const int W = 32 * E;
  for (int i = threadIdx.x; i < 32 * W; i += 256) {
    const int ku = i / W, r = i - ku * W;
    const size_t p = p0 + r / E;
    tile[r / E][ku * E + (r % E)] = (p < P && k0 + ku < Kc) ? in[(((size_t)t * Kc + k0 + ku) * P + p) * E + (r % E)] * mul : 0.f;
  }
  __syncthreads();
  bool bad = false;
  for (int i = threadIdx.x; i < 32 * W; i += 256) {
    const int pp = i / W, e = i - pp * W;
    const size_t p = p0 + pp;
    if (p < P && k0 + e / E < Kc) {
      const float v = tile[pp][e];
      bad |= !(fabsf(v) <= 65504.f);
      out[((size_t)t * P + p) * Kc * E + (size_t)k0 * E + e] = __float2half_rn(v);
    }
  }
  if (bad && overflow) *overflow = 1;
}

// L1 (lama_util.py:12-80) + the head of the script's forward: u8 image [h,w,3] (channel order untouched) and u8 mask
// [h,w] -> NHWC fp16 [H,W,cp]: channels 0..2 = img/255 * (1 - m), channel 3 = m, m = mask > 0; rows / columns beyond
// h, w are the symmetric padding of pad_img_to_modulo (edge sample repeated first).
__global__ void __launch_bounds__(256) rt_lama_input_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, int h, int w,
                                                            __half* __restrict__ out, int H, int W, int cp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int sy = y < h ? y : 2 * h - 1 - y, sx = x < w ? x : 2 * w - 1 - x;
  const float m = mask[(size_t)sy * w + sx] > 0 ? 1.f : 0.f;
  __half* o = out + ((size_t)y * W + x) * cp;
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = __float2half_rn(__fmul_rn(__fdiv_rn((float)img[((size_t)sy * w + sx) * 3 + c], 255.f), 1.f - m));
  o[3] = __float2half_rn(m);
}

// the tail of the script's forward + lama_inpaint.py:25-27: result = m*pred + (1-m)*img/255 (separate fp32 ops, no
// contraction, like torch), u8 = trunc(clip(result*255, 0, 255)), cropped to [h,w].
__global__ void __launch_bounds__(256) rt_lama_output_kernel(const __half* __restrict__ pred, int W, int cp, float inv_scale,
                                                             const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, int h, int w,
                                                             uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const float m = mask[(size_t)y * w + x] > 0 ? 1.f : 0.f;
  const __half* p = pred + ((size_t)y * W + x) * cp;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float im = __fdiv_rn((float)img[((size_t)y * w + x) * 3 + c], 255.f);
    const float r = __fadd_rn(__fmul_rn(m, __half2float(p[c]) * inv_scale), __fmul_rn(1.f - m, im));
    out[((size_t)y * w + x) * 3 + c] = (uint8_t)fminf(fmaxf(__fmul_rn(r, 255.f), 0.f), 255.f);
  }
}

// one channel of an NHWC fp16 tensor -> dense fp32 [pixels], divided by the tensor scale (the probability map leaves the
// device as 4 bytes per pixel instead of the whole 64-channel pitch)
__global__ void __launch_bounds__(256) rt_extract_channel_kernel(const __half* __restrict__ x, size_t pixels, int cp, int ch, float mul,
                                                                 float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < pixels) out[i] = __half2float(x[i * cp + ch]) * mul;
}

// ---- PP-OCRv5 mobile detector (PPLCNetV3 + RSEFPN) ------------------------------------------------------------------
// hardswish followed by the learnable scalar affine of the exported blocks: out = a * hswish(x * inv_s) + c with
// hswish(v) = v * clamp(v + 3, 0, 6) / 6 (not homogeneous: the input scale is divided out first; a and c carry the output scale)
__global__ void __launch_bounds__(256) rt_hswish_affine_kernel(const __half* __restrict__ in, __half* __restrict__ out, size_t n8, float inv_s, float a,
                                                               float c, int* overflow) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = reinterpret_cast<const uint4*>(in)[i];
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  __align__(16) __half2 o[4];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = __half22float2(h[j]);
    f.x *= inv_s;
    f.y *= inv_s;
    const float u = fmaf(a, f.x * fminf(fmaxf(f.x + 3.f, 0.f), 6.f) * (1.f / 6.f), c);
    const float w = fmaf(a, f.y * fminf(fmaxf(f.y + 3.f, 0.f), 6.f) * (1.f / 6.f), c);
    bad |= !(fabsf(u) <= 65504.f) | !(fabsf(w) <= 65504.f);
    o[j] = __floats2half2_rn(u, w);
  }
  if (bad && overflow) *overflow = 1;
  reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(o);
}

// squeeze-and-excitation gate: (1) per-channel mean over the pixels (one block per 8 channels), (2) gate = hardsigmoid(W2 *
// relu(W1 * mean + b1) + b2) in one block; residual != 0 stores 1 + gate (RSELayer: x + x * gate).  The gate then multiplies the
// tensor through the per-channel affine operator.
__global__ void __launch_bounds__(256) rt_channel_mean_kernel(const __half* __restrict__ x, size_t pixels, int cp, float inv_count,
                                                              float* __restrict__ mean) {
  __shared__ float part[8][8];
  const int c8 = blockIdx.x;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (size_t p = threadIdx.x; p < pixels; p += blockDim.x) {
    const uint4 v = *reinterpret_cast<const uint4*>(x + p * cp + c8 * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      acc[2 * j] += f.x;
      acc[2 * j + 1] += f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    for (int o = 16; o; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int j = 0; j < 8; ++j) part[warp][j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
    mean[c8 * 8 + threadIdx.x] = s * inv_count;
  }
}
__global__ void __launch_bounds__(256) rt_se_fc_kernel(const float* __restrict__ mean, const float* __restrict__ w1, const float* __restrict__ b1,
                                                       const float* __restrict__ w2, const float* __restrict__ b2, int C, int mid, float slope,
                                                       float offset, float inv_scale, int residual, float* __restrict__ gate) {
  extern __shared__ float hid[];
  for (int m = threadIdx.x; m < mid; m += blockDim.x) {
    float s = b1[m];
    for (int c = 0; c < C; ++c) s = fmaf(w1[(size_t)m * C + c], mean[c] * inv_scale, s);
    hid[m] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = b2[c];
    for (int m = 0; m < mid; ++m) s = fmaf(w2[(size_t)c * mid + m], hid[m], s);
    const float g = fminf(fmaxf(fmaf(slope, s, offset), 0.f), 1.f);
    gate[c] = residual ? 1.f + g : g;
  }
}

// max |x| over a tensor (calibration of the per-tensor scales): *out is the bit pattern of a non-negative float, so
// an unsigned atomicMax orders it; inf / NaN patterns compare above every finite value.
__global__ void __launch_bounds__(256) rt_absmax_kernel(const __half* __restrict__ x, size_t n8, unsigned int* __restrict__ out) {
  float m = 0.f;
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = reinterpret_cast<const uint4*>(x)[i];
    const __half2* pv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(pv[j]);
      bad |= !(fabsf(f.x) <= 65504.f) | !(fabsf(f.y) <= 65504.f);
      m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  if (bad) m = __int_as_float(0x7f800000);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

// nearest_interp (align_corners False) by an integer factor: out[y][x] = in[y / s][x / s].
__global__ void __launch_bounds__(256) rt_upsample_nearest_kernel(const __half* __restrict__ in, int T, int h, int w, int cp, int s,
                                                                  __half* __restrict__ out, int out_pitch, int out_coff) {
  const int c8n = cp >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * h * s * w * s * c8n;
  if (idx >= total) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = r % (w * s);
  r /= (w * s);
  const int oy = r % (h * s);
  const int t = r / (h * s);
  const uint4 v = *reinterpret_cast<const uint4*>(in + (((size_t)t * h + oy / s) * w + ox / s) * cp + c8 * 8);
  *reinterpret_cast<uint4*>(out + (((size_t)t * h * s + oy) * w * s + ox) * out_pitch + out_coff + c8 * 8) = v;
}

// pool2d max 2x2 stride 1, SAME (pad bottom/right with -inf): out[y][x] = max over y..y+1, x..x+1 inside the image.
__global__ void __launch_bounds__(256) rt_maxpool2x2s1_kernel(const __half* __restrict__ in, int T, int h, int w, int cp,
                                                              __half* __restrict__ out) {
  const int c8n = cp >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * h * w * c8n;
  if (idx >= total) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int x = r % w;
  r /= w;
  const int y = r % h;
  const int t = r / h;
  const __half* base = in + (size_t)t * h * w * cp + c8 * 8;
  uint4 acc = *reinterpret_cast<const uint4*>(base + ((size_t)y * w + x) * cp);
  __half2* pa = reinterpret_cast<__half2*>(&acc);
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      if ((dy | dx) == 0 || y + dy >= h || x + dx >= w) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(base + ((size_t)(y + dy) * w + x + dx) * cp);
      const __half2* pv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) pa[j] = __hmax2(pa[j], pv[j]);
    }
  *reinterpret_cast<uint4*>(out + idx * 8) = acc;
}

// copy C (multiple of 8) channels of every pixel into a channel slice of another tensor (concat)
__global__ void __launch_bounds__(256) rt_copy_channels_kernel(const __half* __restrict__ src, int src_pitch, __half* __restrict__ dst,
                                                               int dst_pitch, int dst_coff, int c8n, size_t pixels) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= pixels * c8n) return;
  const int c8 = idx % c8n;
  const size_t p = idx / c8n;
  *reinterpret_cast<uint4*>(dst + p * dst_pitch + dst_coff + c8 * 8) = *reinterpret_cast<const uint4*>(src + p * src_pitch + c8 * 8);
}

// [T,H,W,cp] -> [T,H/2,W/2,4*cp], channel = ((y&1)*2 + (x&1))*cp + c : input layout of the stride-2 convs
__global__ void __launch_bounds__(256) rt_space_to_depth_kernel(const __half* __restrict__ in, int T, int H, int W, int cp,
                                                                __half* __restrict__ out) {
  const int c8n = cp >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * H * W * c8n;
  if (idx >= total) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int x = r % W;
  r /= W;
  const int y = r % H;
  const int t = r / H;
  const uint4 v = *reinterpret_cast<const uint4*>(in + idx * 8);
  const size_t opix = ((size_t)t * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1);
  *reinterpret_cast<uint4*>(out + opix * 4 * cp + ((y & 1) * 2 + (x & 1)) * cp + c8 * 8) = v;
}

// depthwise conv (groups == channels), any odd kernel, stride 1 or 2, zero padding; weights [k*k][cp] fp32,
// bias [cp]; 8 channels per thread, fp32 accumulation; optional ReLU.
__global__ void __launch_bounds__(256) rt_depthwise_kernel(const __half* __restrict__ in, int T, int H, int W, int cp,
                                                           const float* __restrict__ wgt, const float* __restrict__ bias, int k, int stride,
                                                           int pad, int relu, __half* __restrict__ out, int OH, int OW, RtScale sc) {
  const int c8n = cp >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * OH * OW * c8n;
  if (idx >= total) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = r % OW;
  r /= OW;
  const int oy = r % OH;
  const int t = r / OH;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int ky = 0; ky < k; ++ky) {
    const int iy = oy * stride + ky - pad;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < k; ++kx) {
      const int ix = ox * stride + kx - pad;
      if (ix < 0 || ix >= W) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(in + (((size_t)t * H + iy) * W + ix) * cp + c8 * 8);
      const __half2* pv = reinterpret_cast<const __half2*>(&v);
      const float* wr = wgt + (size_t)(ky * k + kx) * cp + c8 * 8;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(pv[j]);
        acc[2 * j] = fmaf(f.x, wr[2 * j], acc[2 * j]);
        acc[2 * j + 1] = fmaf(f.y, wr[2 * j + 1], acc[2 * j + 1]);
      }
    }
  }
  rt_store8(acc, bias + c8 * 8, sc, relu, out + idx * 8);
}

// Direct conv for tiny channel counts on either side (the 3-channel stem, the 64->1 head): one thread per
// (output pixel, 8 output channels); in [T,H,W,cin_p] fp16, weights [k*k*cin][cout_p] fp32 (cout_p multiple of 8).
__global__ void __launch_bounds__(256) rt_direct_conv_kernel(const __half* __restrict__ in, int T, int H, int W, int cin_p, int cin,
                                                             const float* __restrict__ wgt, const float* __restrict__ bias, int kh, int kw,
                                                             int stride, int pad_t, int pad_l, int relu, __half* __restrict__ out, int OH,
                                                             int OW, int cout_p, int out_pitch, RtScale sc) {
  const int c8n = cout_p >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * OH * OW * c8n;
  if (idx >= total) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = r % OW;
  r /= OW;
  const int oy = r % OH;
  const int t = r / OH;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int ky = 0; ky < kh; ++ky) {
    const int iy = oy * stride + ky - pad_t;
    if (iy < 0 || iy >= H) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int ix = ox * stride + kx - pad_l;
      if (ix < 0 || ix >= W) continue;
      const __half* px = in + (((size_t)t * H + iy) * W + ix) * cin_p;
      const float* wr = wgt + (size_t)((ky * kw + kx) * cin) * cout_p + c8 * 8;
      for (int ci = 0; ci < cin; ++ci) {
        const float v = __half2float(px[ci]);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[(size_t)ci * cout_p + j], acc[j]);
      }
    }
  }
  rt_store8(acc, bias + c8 * 8, sc, relu, out + (idx / c8n) * out_pitch + c8 * 8);
}

// The same direct conv with its weights in shared memory and two horizontally adjacent output pixels per thread: the kernel above reads
// every weight from global memory for every FMA group (one L1 load per FMA), which made the 5 -> 64 first layer of the ProPainter encoder
// and RAFT's 7x7 3 -> 64 stem cost milliseconds at full resolution (profiles/launches_r2_propainter.md: 32 % of the pipeline's kernel
// time).  Here a block stages the kh*kw*cin*cout_p weights once (dynamic smem, fp32), a thread keeps 2 x 8 accumulators and feeds 16
// FMAs from two 16-byte shared loads (warp lanes with the same channel group broadcast).  Same arithmetic order per output as above.
__global__ void __launch_bounds__(256) rt_direct_conv_smem_kernel(const __half* __restrict__ in, int T, int H, int W, int cin_p, int cin,
                                                                  const float* __restrict__ wgt, const float* __restrict__ bias, int kh, int kw,
                                                                  int stride, int pad_t, int pad_l, int relu, __half* __restrict__ out, int OH,
                                                                  int OW, int cout_p, int out_pitch, RtScale sc) {
  extern __shared__ __align__(16) float sw[];
  const int nw = kh * kw * cin * cout_p;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = wgt[i];
  __syncthreads();
  const int c8n = cout_p >> 3, OWp = (OW + 1) >> 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * OH * OWp * c8n;
  if (idx >= total) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = (int)(r % OWp) * 2;
  r /= OWp;
  const int oy = r % OH;
  const int t = r / OH;
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  for (int ky = 0; ky < kh; ++ky) {
    const int iy = oy * stride + ky - pad_t;
    if (iy < 0 || iy >= H) continue;
    const __half* row = in + ((size_t)t * H + iy) * W * cin_p;
    for (int kx = 0; kx < kw; ++kx) {
      const int ix0 = ox * stride + kx - pad_l, ix1 = ix0 + stride;
      const bool ok0 = ix0 >= 0 && ix0 < W, ok1 = ix1 >= 0 && ix1 < W;
      if (!ok0 && !ok1) continue;
      const __half* p0 = row + (size_t)(ok0 ? ix0 : 0) * cin_p;
      const __half* p1 = row + (size_t)(ok1 ? ix1 : 0) * cin_p;
      const float* wr = sw + (size_t)((ky * kw + kx) * cin) * cout_p + c8 * 8;
      for (int ci = 0; ci < cin; ++ci) {
        const float v0 = ok0 ? __half2float(p0[ci]) : 0.f, v1 = ok1 ? __half2float(p1[ci]) : 0.f;
        const float4 w0 = *reinterpret_cast<const float4*>(wr + (size_t)ci * cout_p);
        const float4 w1 = *reinterpret_cast<const float4*>(wr + (size_t)ci * cout_p + 4);
        a0[0] = fmaf(v0, w0.x, a0[0]); a0[1] = fmaf(v0, w0.y, a0[1]); a0[2] = fmaf(v0, w0.z, a0[2]); a0[3] = fmaf(v0, w0.w, a0[3]);
        a0[4] = fmaf(v0, w1.x, a0[4]); a0[5] = fmaf(v0, w1.y, a0[5]); a0[6] = fmaf(v0, w1.z, a0[6]); a0[7] = fmaf(v0, w1.w, a0[7]);
        a1[0] = fmaf(v1, w0.x, a1[0]); a1[1] = fmaf(v1, w0.y, a1[1]); a1[2] = fmaf(v1, w0.z, a1[2]); a1[3] = fmaf(v1, w0.w, a1[3]);
        a1[4] = fmaf(v1, w1.x, a1[4]); a1[5] = fmaf(v1, w1.y, a1[5]); a1[6] = fmaf(v1, w1.z, a1[6]); a1[7] = fmaf(v1, w1.w, a1[7]);
      }
    }
  }
  const size_t opix = ((size_t)t * OH + oy) * OW + ox;
  rt_store8(a0, bias + c8 * 8, sc, relu, out + opix * out_pitch + c8 * 8);
  if (ox + 1 < OW) rt_store8(a1, bias + c8 * 8, sc, relu, out + (opix + 1) * out_pitch + c8 * 8);
}

// conv2d_transpose 2x2 stride 2 (no overlap): out[2y+dy][2x+dx][co] = bias[co] + sum_ci in[y][x][ci] * w[ci][co][dy][dx].
// weights repacked as [dy*2+dx][cin][cout_p] fp32.  One thread per (output pixel, 8 output channels).
__global__ void __launch_bounds__(256) rt_deconv2x2_kernel(const __half* __restrict__ in, int T, int H, int W, int cin_p, int cin,
                                                           const float* __restrict__ wgt, const float* __restrict__ bias, int relu,
                                                           __half* __restrict__ out, int cout_p, int out_pitch, RtScale sc) {
  const int c8n = cout_p >> 3;
  const int OH = 2 * H, OW = 2 * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)T * OH * OW * c8n;
  if (idx >= total) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = r % OW;
  r /= OW;
  const int oy = r % OH;
  const int t = r / OH;
  const __half* px = in + (((size_t)t * H + (oy >> 1)) * W + (ox >> 1)) * cin_p;
  const float* wr = wgt + (size_t)(((oy & 1) * 2 + (ox & 1)) * cin) * cout_p + c8 * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int ci = 0; ci < cin; ++ci) {
    const float v = __half2float(px[ci]);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wr[(size_t)ci * cout_p + j], acc[j]);
  }
  rt_store8(acc, bias + c8 * 8, sc, relu, out + (idx / c8n) * out_pitch + c8 * 8);
}

// DetResizeForTest + NormalizeImage + ToCHW of the detector (inference.yml:22-40): BGR u8 [sh,sw,3] ->
// cv2-exact bilinear resize -> (x/255 - mean[c]) / std[c] in BGR order -> NHWC fp16 [dh,dw,cp] (3 real channels).
struct ResizeTaps;
__global__ void __launch_bounds__(256) rt_det_preprocess_kernel(const uint8_t* __restrict__ src, int sw, int sh, __half* __restrict__ dst,
                                                                int dw, int dh, int cp, const int* __restrict__ xi0,
                                                                const int* __restrict__ xi1, const short* __restrict__ xw0,
                                                                const short* __restrict__ xw1, const int* __restrict__ yi0,
                                                                const int* __restrict__ yi1, const short* __restrict__ yw0,
                                                                const short* __restrict__ yw1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= dw) return;
  const uint8_t* r0 = src + (size_t)yi0[y] * sw * 3;
  const uint8_t* r1 = src + (size_t)yi1[y] * sw * 3;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  __half* o = dst + ((size_t)y * dw + x) * cp;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int s0 = r0[xi0[x] * 3 + c] * xw0[x] + r0[xi1[x] * 3 + c] * xw1[x];
    const int s1 = r1[xi0[x] * 3 + c] * xw0[x] + r1[xi1[x] * 3 + c] * xw1[x];
    int v = (((yw0[y] * (s0 >> 4)) >> 16) + ((yw1[y] * (s1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    o[c] = __float2half_rn(((float)v * (1.0f / 255.0f) - mean[c]) / stdv[c]);
  }
}

}  // namespace vsr
