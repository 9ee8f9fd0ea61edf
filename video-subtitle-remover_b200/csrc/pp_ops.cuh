// RAFT operators of the ProPainter path (backend/inpaint/video/raft/*.py) for the graph runtime.
// STATUS: written and checked against a CPU stand-in of the runtime (tests/fake_rt.py, tests/test_raft_cpu.py); NOT yet run
// on a B200 — round 1's GPU budget was spent before this file existed.  Nothing else in the engine uses it.
//
// Layouts: NHWC fp16 tensors [N,h,w,pitch]; the flow state keeps an fp32 master [N*h*w][2] next to its fp16 copy [N*h*w][8];
// correlation volumes are [N][h*w source pixels][pitch >= h2*w2 target pixels] fp16, one buffer per pyramid level.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vsr {

// frames u8 [P][3] (BGR as cv2 gives them) -> fp16 [P][8]: RGB order, x/255*2-1 (propainter_inpaint.py:194,216 + to_tensors())
__global__ void __launch_bounds__(256) pp_frames_to_half_kernel(const uint8_t* __restrict__ bgr, size_t pixels, __half* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  __align__(16) __half o[8];
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = __float2half_rn(__fdiv_rn((float)bgr[i * 3 + (2 - c)], 255.f) * 2.f - 1.f);
#pragma unroll
  for (int c = 3; c < 8; ++c) o[c] = __float2half_rn(0.f);
  *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(o);
}

// nn.InstanceNorm2d (no affine, eps 1e-5, biased variance): statistics per (image, channel); one block = 8 channels of one image
__global__ void __launch_bounds__(256) pp_instnorm_stats_kernel(const __half* __restrict__ x, size_t pixels, int cp, float* __restrict__ mean,
                                                                float* __restrict__ rstd) {
  __shared__ float part[8][16];
  const int n = blockIdx.y, c8 = blockIdx.x;
  const __half* base = x + (size_t)n * pixels * cp + c8 * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  for (size_t p = threadIdx.x; p < pixels; p += blockDim.x) {
    const uint4 v = *reinterpret_cast<const uint4*>(base + p * cp);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h[j]);
      s[2 * j] += f.x; q[2 * j] += f.x * f.x;
      s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    for (int o = 16; o; o >>= 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
      q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
    }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int j = 0; j < 8; ++j) { part[warp][j] = s[j]; part[warp][8 + j] = q[j]; }
  __syncthreads();
  if (threadIdx.x < 8) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += part[w][threadIdx.x]; b += part[w][8 + threadIdx.x]; }
    const float m = a / (float)pixels;
    const float var = fmaxf(b / (float)pixels - m * m, 0.f);
    mean[(size_t)n * cp + c8 * 8 + threadIdx.x] = m;
    rstd[(size_t)n * cp + c8 * 8 + threadIdx.x] = rsqrtf(var + 1e-5f);
  }
}
__global__ void __launch_bounds__(256) pp_instnorm_apply_kernel(const __half* __restrict__ x, size_t pixels, int cp, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, int relu, __half* __restrict__ out, size_t total8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8n = cp >> 3;
  const int c0 = (int)(i % c8n) * 8;
  const size_t n = i / ((size_t)pixels * c8n);
  const uint4 v = reinterpret_cast<const uint4*>(x)[i];
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  const float* m = mean + n * cp + c0;
  const float* r = rstd + n * cp + c0;
  __align__(16) __half2 o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    float a = (f.x - m[2 * j]) * r[2 * j], b = (f.y - m[2 * j + 1]) * r[2 * j + 1];
    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
    o[j] = __floats2half2_rn(a, b);
  }
  reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(o);
}

// cnet output [P][256] -> net = tanh(first 128) into a slice of pitch pn, inp = relu(last 128) into a slice of pitch pi (raft.py:116-118)
__global__ void __launch_bounds__(256) pp_context_split_kernel(const __half* __restrict__ x, size_t pixels, __half* __restrict__ net, int pn,
                                                               __half* __restrict__ inp, int pi) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * 32) return;
  const int c8 = i % 32;
  const size_t p = i / 32;
  const uint4 v = *reinterpret_cast<const uint4*>(x + p * 256 + c8 * 8);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  __align__(16) __half2 o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(h[j]);
    o[j] = c8 < 16 ? __floats2half2_rn(tanhf(f.x), tanhf(f.y)) : __floats2half2_rn(fmaxf(f.x, 0.f), fmaxf(f.y, 0.f));
  }
  __half* dst = c8 < 16 ? net + p * pn + c8 * 8 : inp + p * pi + (c8 - 16) * 8;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(o);
}

// F.avg_pool2d(corr, 2, stride=2) over the TARGET dims of a correlation volume: in [rows][pitch_in] viewed as [h2][w2] per row
__global__ void __launch_bounds__(256) pp_corr_pool_kernel(const __half* __restrict__ in, size_t rows, int h2, int w2, int pitch_in,
                                                           __half* __restrict__ out, int pitch_out) {
  const int oh = h2 >> 1, ow = w2 >> 1;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * oh * ow) return;
  const int x = i % ow;
  const int y = (i / ow) % oh;
  const size_t r = i / ((size_t)ow * oh);
  const __half* s = in + r * pitch_in + (size_t)(2 * y) * w2 + 2 * x;
  const float v = (__half2float(s[0]) + __half2float(s[1]) + __half2float(s[w2]) + __half2float(s[w2 + 1])) * 0.25f;
  out[r * pitch_out + (size_t)y * ow + x] = __float2half_rn(v);
}

struct CorrLevels {
  const __half* ptr[4];
  int h[4], w[4], pitch[4];
};
// CorrBlock.__call__ (corr.py:29-50): for source pixel p of pair n at coords1 = (x + flow.x, y + flow.y): 4 levels x 9x9 bilinear
// samples (zeros outside, align_corners=True) of its correlation map at coords1 / 2^l + delta.  The reference adds the FIRST
// meshgrid component (the row offset list) to x: sample (i, j) sits at (x + d[i], y + d[j]); kept.  out [P][pitch], channel l*81 + i*9 + j.
__global__ void __launch_bounds__(256) pp_corr_lookup_kernel(const __grid_constant__ CorrLevels lv, const float* __restrict__ flow, int h, int w, size_t pixels,
                                                             __half* __restrict__ out, int pitch) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * 324) return;
  const int k = i % 324;
  const size_t p = i / 324;
  const int l = k / 81, ij = k - l * 81, ii = ij / 9, jj = ij - ii * 9;
  const int px = (int)(p % w), py = (int)((p / w) % h);
  const float scale = 1.f / (float)(1 << l);
  const float x = ((float)px + flow[p * 2]) * scale + (float)(ii - 4);
  const float y = ((float)py + flow[p * 2 + 1]) * scale + (float)(jj - 4);
  const int H = lv.h[l], W = lv.w[l];
  const __half* m = lv.ptr[l] + p * lv.pitch[l];
  const float fx = floorf(x), fy = floorf(y);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = x - fx, ay = y - fy;
  float v = 0.f;
  if (y0 >= 0 && y0 < H) {
    if (x0 >= 0 && x0 < W) v += (1.f - ax) * (1.f - ay) * __half2float(m[(size_t)y0 * W + x0]);
    if (x0 + 1 >= 0 && x0 + 1 < W) v += ax * (1.f - ay) * __half2float(m[(size_t)y0 * W + x0 + 1]);
  }
  if (y0 + 1 >= 0 && y0 + 1 < H) {
    if (x0 >= 0 && x0 < W) v += (1.f - ax) * ay * __half2float(m[(size_t)(y0 + 1) * W + x0]);
    if (x0 + 1 >= 0 && x0 + 1 < W) v += ax * ay * __half2float(m[(size_t)(y0 + 1) * W + x0 + 1]);
  }
  out[p * pitch + k] = __float2half_rn(v);
}

// SepConvGRU (update.py:34-57), one half step: rh = sigmoid(r) * h (into the q-conv input), then h = (1 - sigmoid(z)) * h + sigmoid(z) * tanh(q)
__global__ void __launch_bounds__(256) pp_gru_rh_kernel(const __half* __restrict__ r, int pr, const __half* __restrict__ hsrc, int ph,
                                                        __half* __restrict__ out, int po, size_t pixels) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * 16) return;
  const int c8 = i % 16;
  const size_t p = i / 16;
  const uint4 vr = *reinterpret_cast<const uint4*>(r + p * pr + c8 * 8), vh = *reinterpret_cast<const uint4*>(hsrc + p * ph + c8 * 8);
  const __half2* a = reinterpret_cast<const __half2*>(&vr);
  const __half2* b = reinterpret_cast<const __half2*>(&vh);
  __align__(16) __half2 o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fr = __half22float2(a[j]), fh = __half22float2(b[j]);
    o[j] = __floats2half2_rn(fh.x / (1.f + __expf(-fr.x)), fh.y / (1.f + __expf(-fr.y)));
  }
  *reinterpret_cast<uint4*>(out + p * po + c8 * 8) = *reinterpret_cast<const uint4*>(o);
}
__global__ void __launch_bounds__(256) pp_gru_update_kernel(const __half* __restrict__ z, int pz, const __half* __restrict__ q, int pq,
                                                            __half* __restrict__ hio, int ph, size_t pixels) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * 16) return;
  const int c8 = i % 16;
  const size_t p = i / 16;
  const uint4 vz = *reinterpret_cast<const uint4*>(z + p * pz + c8 * 8), vq = *reinterpret_cast<const uint4*>(q + p * pq + c8 * 8);
  uint4 vh = *reinterpret_cast<const uint4*>(hio + p * ph + c8 * 8);
  const __half2* a = reinterpret_cast<const __half2*>(&vz);
  const __half2* b = reinterpret_cast<const __half2*>(&vq);
  __half2* c = reinterpret_cast<__half2*>(&vh);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fz = __half22float2(a[j]), fq = __half22float2(b[j]), fh = __half22float2(c[j]);
    const float zx = 1.f / (1.f + __expf(-fz.x)), zy = 1.f / (1.f + __expf(-fz.y));
    c[j] = __floats2half2_rn((1.f - zx) * fh.x + zx * tanhf(fq.x), (1.f - zy) * fh.y + zy * tanhf(fq.y));
  }
  *reinterpret_cast<uint4*>(hio + p * ph + c8 * 8) = vh;
}

// flow32 += delta (the flow head's 2 channels, fp16 [P][pd]); refresh the fp16 copies: flow16 [P][8] (the motion encoder's input) and
// the last two channels of the motion features inside the GRU input tensors (torch.cat([out, flow]), update.py:96)
__global__ void __launch_bounds__(256) pp_flow_update_kernel(float* __restrict__ flow32, const __half* __restrict__ delta, int pd,
                                                             __half* __restrict__ flow16, __half* __restrict__ dst_a, __half* __restrict__ dst_b,
                                                             int pitch_ab, int coff, size_t pixels, int add) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  float fx = flow32[p * 2], fy = flow32[p * 2 + 1];
  if (add) {
    fx += __half2float(delta[p * pd]);
    fy += __half2float(delta[p * pd + 1]);
    flow32[p * 2] = fx;
    flow32[p * 2 + 1] = fy;
  }
  const __half hx = __float2half_rn(fx), hy = __float2half_rn(fy);
  __align__(16) __half o[8] = {hx, hy, __half(), __half(), __half(), __half(), __half(), __half()};
  *reinterpret_cast<uint4*>(flow16 + p * 8) = *reinterpret_cast<const uint4*>(o);
  if (dst_a) { dst_a[p * pitch_ab + coff] = hx; dst_a[p * pitch_ab + coff + 1] = hy; }
  if (dst_b) { dst_b[p * pitch_ab + coff] = hx; dst_b[p * pitch_ab + coff + 1] = hy; }
}

// RAFT.upsample_flow (raft.py:73-84): out[n][c][8y+sy][8x+sx] = sum_k softmax_k(mask[n][y][x][k*64 + sy*8 + sx]) * 8 * flow[n][y+ky-1][x+kx-1][c]
// (3x3 neighbourhood with zero padding, k = ky*3 + kx); out is planar fp32 [N][2][8h][8w]
__global__ void __launch_bounds__(256) pp_convex_upsample_kernel(const float* __restrict__ flow32, const __half* __restrict__ mask, int pm, int N, int h,
                                                                 int w, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)N * h * w * 64;
  if (i >= total) return;
  const int s = i % 64, sy = s >> 3, sx = s & 7;
  const size_t p = i / 64;
  const int x = (int)(p % w), y = (int)((p / w) % h);
  const int n = (int)(p / ((size_t)w * h));
  const __half* m = mask + p * pm;
  float e[9], mx = -1e30f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { e[k] = __half2float(m[k * 64 + s]); mx = fmaxf(mx, e[k]); }
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { e[k] = __expf(e[k] - mx); den += e[k]; }
  float ux = 0.f, uy = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
    const float* f = flow32 + (((size_t)n * h + yy) * w + xx) * 2;
    ux += e[k] * f[0];
    uy += e[k] * f[1];
  }
  const float g = 8.f / den;
  const size_t H8 = (size_t)8 * h, W8 = (size_t)8 * w;
  const size_t o = ((size_t)n * 2 * H8 + (size_t)(8 * y + sy)) * W8 + (size_t)(8 * x + sx);
  out[o] = ux * g;
  out[o + H8 * W8] = uy * g;
}

// ---- P5: InpaintGenerator.img_propagation = BidirectionalPropagation(learnable=False) (propainter.py:107-193) ---------------------
// One step of the sequential propagation for one frame.  State pixels are 8 halves: channels 0..2 the (masked) frame in [-1, 1],
// channel 3 the mask (1 = still missing).  prev = the propagated previous frame of this pass, cur = this frame of the pass input;
// flow_prop / flow_check: planar fp32 [2][H][W] (x then y).  Per pixel (all of :160-176):
//   valid = |f + bilinear(flow_check)(p + f)|^2 < 0.01 (|f|^2 + |.|^2) + 0.5            fbConsistencyCheck :24-33
//   warped = nearest(prev frame)(p + f), hole_there = bilinear(prev mask)(p + f) > 0.1    flow_warp(.., 'nearest') / default bilinear
//   use = cur mask & valid & !hole_there;  out frame = use ? warped : cur;  out mask = cur mask & !(valid & !hole_there)
// grid_sample semantics: align_corners=True (pixel coordinates), zeros outside, nearest = round half to even.
__device__ __forceinline__ float pp_bilinear_plane(const float* __restrict__ m, int H, int W, float x, float y) {
  const float fx = floorf(x), fy = floorf(y);
  const int x0 = (int)fx, y0 = (int)fy;
  const float ax = x - fx, ay = y - fy;
  float v = 0.f;
  if (y0 >= 0 && y0 < H) {
    if (x0 >= 0 && x0 < W) v += (1.f - ax) * (1.f - ay) * m[(size_t)y0 * W + x0];
    if (x0 + 1 >= 0 && x0 + 1 < W) v += ax * (1.f - ay) * m[(size_t)y0 * W + x0 + 1];
  }
  if (y0 + 1 >= 0 && y0 + 1 < H) {
    if (x0 >= 0 && x0 < W) v += (1.f - ax) * ay * m[(size_t)(y0 + 1) * W + x0];
    if (x0 + 1 >= 0 && x0 + 1 < W) v += ax * ay * m[(size_t)(y0 + 1) * W + x0 + 1];
  }
  return v;
}
__global__ void __launch_bounds__(256) pp_img_prop_step_kernel(const __half* __restrict__ prev, const __half* __restrict__ cur, const float* __restrict__ flow_prop,
                                                               const float* __restrict__ flow_check, int H, int W, __half* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x, plane = (size_t)H * W;
  const float fx = flow_prop[p], fy = flow_prop[plane + p];
  const float sx = (float)x + fx, sy = (float)y + fy;
  const float bx = pp_bilinear_plane(flow_check, H, W, sx, sy), by = pp_bilinear_plane(flow_check + plane, H, W, sx, sy);
  const float dx = fx + bx, dy = fy + by;
  const bool valid = dx * dx + dy * dy < 0.01f * (fx * fx + fy * fy + bx * bx + by * by) + 0.5f;
  // bilinear sample of the previous mask (channel 3 of the state)
  float hole = 0.f;
  {
    const float gx = floorf(sx), gy = floorf(sy);
    const int x0 = (int)gx, y0 = (int)gy;
    const float ax = sx - gx, ay = sy - gy;
    auto mk = [&](int yy, int xx) { return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __half2float(prev[((size_t)yy * W + xx) * 8 + 3]) : 0.f; };
    hole = (1.f - ax) * (1.f - ay) * mk(y0, x0) + ax * (1.f - ay) * mk(y0, x0 + 1) + (1.f - ax) * ay * mk(y0 + 1, x0) + ax * ay * mk(y0 + 1, x0 + 1);
  }
  const bool hole_there = hole > 0.1f;
  uint4 c = *reinterpret_cast<const uint4*>(cur + p * 8);
  __half* ch = reinterpret_cast<__half*>(&c);
  const bool cur_hole = __half2float(ch[3]) > 0.1f;
  const bool fill = valid && !hole_there;
  if (cur_hole && fill) {
    const int nx = (int)nearbyintf(sx), ny = (int)nearbyintf(sy);      // nearest, round half to even
    float w3[3] = {0.f, 0.f, 0.f};
    if (nx >= 0 && nx < W && ny >= 0 && ny < H) {
      const __half* q = prev + ((size_t)ny * W + nx) * 8;
      w3[0] = __half2float(q[0]); w3[1] = __half2float(q[1]); w3[2] = __half2float(q[2]);
    }
    ch[0] = __float2half_rn(w3[0]); ch[1] = __float2half_rn(w3[1]); ch[2] = __float2half_rn(w3[2]);
  }
  ch[3] = __float2half_rn((cur_hole && !fill) ? 1.f : 0.f);
  *reinterpret_cast<uint4*>(out + p * 8) = c;
}

// state pixels from frames and masks: frame * (1 - m) | m  (propainter_inpaint.py:283), frames fp16 [P][8] RGB in [-1,1], mask u8 [H][W] shared
// by all frames (> 0 = hole) ; and the inverse composition  updated = frame * (1 - m) + prop * m  (:311) into channels 0..2 (+ updated mask in 3)
__global__ void __launch_bounds__(256) pp_state_init_kernel(const __half* __restrict__ frames, const uint8_t* __restrict__ mask, size_t plane, size_t pixels,
                                                            __half* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  uint4 v = *reinterpret_cast<const uint4*>(frames + i * 8);
  __half* h = reinterpret_cast<__half*>(&v);
  const bool m = mask[i % plane] > 0;
  if (m) { h[0] = __float2half_rn(0.f); h[1] = h[0]; h[2] = h[0]; }
  h[3] = __float2half_rn(m ? 1.f : 0.f);
  *reinterpret_cast<uint4*>(out + i * 8) = v;
}
__global__ void __launch_bounds__(256) pp_state_compose_kernel(const __half* __restrict__ frames, const uint8_t* __restrict__ mask, const __half* __restrict__ prop,
                                                               size_t plane, size_t pixels, __half* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  uint4 f = *reinterpret_cast<const uint4*>(frames + i * 8);
  const uint4 q = *reinterpret_cast<const uint4*>(prop + i * 8);
  __half* h = reinterpret_cast<__half*>(&f);
  const __half* g = reinterpret_cast<const __half*>(&q);
  if (mask[i % plane] > 0) { h[0] = g[0]; h[1] = g[1]; h[2] = g[2]; }
  h[3] = g[3];
  *reinterpret_cast<uint4*>(out + i * 8) = f;
}

// ---- P4: RecurrentFlowCompleteNet (video/model/recurrent_flow_completion.py) ----------------------------------------------------
// network input (:305-317): per frame [fx * (1 - m), fy * (1 - m), m] from planar fp32 flows [N][2][H][W] and one u8 mask [H][W];
// reverse != 0 reads the flows in reversed time order (the backward direction runs the same network on the flipped sequence)
__global__ void __launch_bounds__(256) pp_rfc_input_kernel(const float* __restrict__ flow, const uint8_t* __restrict__ mask, int N, size_t plane, int reverse,
                                                           __half* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * plane) return;
  const size_t n = i / plane, p = i % plane;
  const float* f = flow + (size_t)(reverse ? N - 1 - n : n) * 2 * plane;
  const float m = mask[p] > 0 ? 1.f : 0.f;
  __align__(16) __half o[8];
  o[0] = __float2half_rn(f[p] * (1.f - m));
  o[1] = __float2half_rn(f[plane + p] * (1.f - m));
  o[2] = __float2half_rn(m);
#pragma unroll
  for (int c = 3; c < 8; ++c) o[c] = __float2half_rn(0.f);
  *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(o);
}

// padding_mode='replicate' of the first Conv3d (:213-214): out[OH][OW] = in[clamp(y - top)][clamp(x - left)], 8 channels per thread
__global__ void __launch_bounds__(256) pp_pad_replicate_kernel(const __half* __restrict__ in, int T, int H, int W, int cp, __half* __restrict__ out, int OH, int OW,
                                                               int top, int left) {
  const int c8n = cp >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)T * OH * OW * c8n) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int ox = r % OW;
  r /= OW;
  const int oy = r % OH;
  const int t = r / OH;
  const int y = min(max(oy - top, 0), H - 1), x = min(max(ox - left, 0), W - 1);
  *reinterpret_cast<uint4*>(out + idx * 8) = *reinterpret_cast<const uint4*>(in + (((size_t)t * H + y) * W + x) * cp + c8 * 8);
}

// LeakyReLU(slope) in place over a whole tensor
__global__ void __launch_bounds__(256) pp_leaky_relu_kernel(__half* __restrict__ x, size_t n8, float slope) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 v = reinterpret_cast<uint4*>(x)[i];
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = __half22float2(h[j]);
    f.x = f.x > 0.f ? f.x : f.x * slope;
    f.y = f.y > 0.f ? f.y : f.y * slope;
    h[j] = __floats2half2_rn(f.x, f.y);
  }
  reinterpret_cast<uint4*>(x)[i] = v;
}

// temporal taps of P3DBlock.conv2 (Conv3d (3,1,1), dilation (2,1,1), padding (2,0,0), :164-165): out[t][p][k*C + c] = in[t + 2(k-1)][p][c], zero
// outside the clip; the 1x1 conv that follows is the temporal conv.  C = cp_in (multiple of 8), out pitch >= 3*C.
__global__ void __launch_bounds__(256) pp_temporal_taps_kernel(const __half* __restrict__ in, int T, size_t pixels, int cp_in, __half* __restrict__ out, int cp_out) {
  const int c8n = cp_in >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)T * pixels * 3 * c8n) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int k = r % 3;
  r /= 3;
  const size_t p = r % pixels;
  const int t = (int)(r / pixels);
  const int ts = t + 2 * (k - 1);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (ts >= 0 && ts < T) v = *reinterpret_cast<const uint4*>(in + ((size_t)ts * pixels + p) * cp_in + c8 * 8);
  *reinterpret_cast<uint4*>(out + ((size_t)t * pixels + p) * cp_out + (size_t)k * cp_in + c8 * 8) = v;
}

// modulated deformable 3x3 conv, pad 1, stride 1 (torchvision.ops.deform_conv2d as called at :44-46 / propainter.py:69-72), gather half:
// cols[p][k*C + c] = sigmoid(om[p][2*G*9 + g*9 + k]) * bilinear(x[:, c])(y - 1 + ky + dy, x - 1 + kx + dx), g = c / (C / G),
// (dy, dx) = max_res * tanh(om[p][2*(g*9+k) + {0,1}]) (+ (flow.y, flow.x) when flow != nullptr, propainter.py:64).  x = [xa | xb] channel-wise
// (two sources of Ca and C - Ca channels; xb may be nullptr when Ca == C).  The 1x1 conv over cols (K = 9*C) is the rest of the operator.
__global__ void __launch_bounds__(256) pp_deform_cols_kernel(const __half* __restrict__ xa, int pa, int Ca, const __half* __restrict__ xb, int pb, int C, int G,
                                                             const __half* __restrict__ om, int pom, float max_res, const float* __restrict__ flow, int H, int W,
                                                             size_t pixels, __half* __restrict__ cols, int pcols) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int groups_k = G * 9;
  if (idx >= pixels * groups_k) return;
  const int gk = idx % groups_k;
  const size_t p = idx / groups_k;
  const int g = gk / 9, k = gk - g * 9, ky = k / 3, kx = k - ky * 3;
  const int px = (int)(p % W), py = (int)((p / W) % H);
  const size_t img = p / ((size_t)W * H) * ((size_t)W * H);
  const __half* o = om + p * pom;
  float dy = max_res * tanhf(__half2float(o[2 * gk])), dx = max_res * tanhf(__half2float(o[2 * gk + 1]));
  if (flow) { dy += flow[p * 2 + 1]; dx += flow[p * 2]; }
  const float m = 1.f / (1.f + __expf(-__half2float(o[2 * groups_k + gk])));
  const float sy = (float)(py - 1 + ky) + dy, sx = (float)(px - 1 + kx) + dx;
  const float fy = floorf(sy), fx = floorf(sx);
  const int y0 = (int)fy, x0 = (int)fx;
  const float ay = sy - fy, ax = sx - fx;
  const float w00 = (1.f - ay) * (1.f - ax) * m, w01 = (1.f - ay) * ax * m, w10 = ay * (1.f - ax) * m, w11 = ay * ax * m;
  const bool v00 = y0 >= 0 && y0 < H && x0 >= 0 && x0 < W, v01 = y0 >= 0 && y0 < H && x0 + 1 >= 0 && x0 + 1 < W;
  const bool v10 = y0 + 1 >= 0 && y0 + 1 < H && x0 >= 0 && x0 < W, v11 = y0 + 1 >= 0 && y0 + 1 < H && x0 + 1 >= 0 && x0 + 1 < W;
  const int cpg = C / G;
  if (cpg == 8 && (Ca & 7) == 0 && (pa & 7) == 0 && (pb & 7) == 0 && (C & 7) == 0 && (pcols & 7) == 0) {
    // the networks' case (128 channels, 16 groups): the 8 channels of a group are one 16-byte load per corner and one 16-byte store
    const int c = g * 8;
    const __half* src = c < Ca ? xa + c : xb + (c - Ca);
    const int pitch = c < Ca ? pa : pb;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto corner = [&](bool ok, float wgt, int yy, int xx) {
      if (!ok) return;
      const uint4 raw = *reinterpret_cast<const uint4*>(src + (img + (size_t)yy * W + xx) * pitch);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        acc[2 * j] += wgt * f.x;
        acc[2 * j + 1] += wgt * f.y;
      }
    };
    corner(v00, w00, y0, x0);
    corner(v01, w01, y0, x0 + 1);
    corner(v10, w10, y0 + 1, x0);
    corner(v11, w11, y0 + 1, x0 + 1);
    __align__(16) __half o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(acc[j]);
    *reinterpret_cast<uint4*>(cols + p * pcols + (size_t)k * C + c) = *reinterpret_cast<const uint4*>(o);
    return;
  }
  for (int cc = 0; cc < cpg; ++cc) {
    const int c = g * cpg + cc;
    const __half* src = c < Ca ? xa + c : xb + (c - Ca);
    const int pitch = c < Ca ? pa : pb;
    float v = 0.f;
    if (v00) v += w00 * __half2float(src[(img + (size_t)y0 * W + x0) * pitch]);
    if (v01) v += w01 * __half2float(src[(img + (size_t)y0 * W + x0 + 1) * pitch]);
    if (v10) v += w10 * __half2float(src[(img + (size_t)(y0 + 1) * W + x0) * pitch]);
    if (v11) v += w11 * __half2float(src[(img + (size_t)(y0 + 1) * W + x0 + 1) * pitch]);
    cols[p * pcols + (size_t)k * C + c] = __float2half_rn(v);
  }
}

// ---- P6: InpaintGenerator.forward (video/model/propainter.py:321-378), front half ---------------------------------------------------
// encoder input (:333-335): frame ids[i] of the propagation state [T][H][W][8] (rgb | updated mask) + the dilated input mask ->
// [rgb, m_in, m_updated, 0, 0, 0] fp16 [n][H][W][8]
__global__ void __launch_bounds__(256) pp_gen_input_kernel(const __half* __restrict__ state, const uint8_t* __restrict__ mask, const int* __restrict__ ids, int n,
                                                           size_t plane, __half* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * plane) return;
  const size_t k = i / plane, p = i % plane;
  uint4 v = *reinterpret_cast<const uint4*>(state + ((size_t)ids[k] * plane + p) * 8);
  __half* h = reinterpret_cast<__half*>(&v);
  h[4] = h[3];
  h[3] = __float2half_rn(mask[p] > 0 ? 1.f : 0.f);
  h[5] = h[6] = h[7] = __float2half_rn(0.f);
  *reinterpret_cast<uint4*>(out + i * 8) = v;
}

// F.interpolate(flow, scale_factor=1/4, mode='bilinear', align_corners=False) / 4 (:341-344): the sample point of output (y, x) is
// (4y + 1.5, 4x + 1.5), i.e. the mean of the 2x2 block at (4y+1, 4x+1).  in planar fp32 [n][2][H][W] (frame ids[i]) -> out fp32 [n*h*w][2]
__global__ void __launch_bounds__(256) pp_flow_down4_kernel(const float* __restrict__ flow, const int* __restrict__ ids, int n, int H, int W,
                                                            float* __restrict__ out) {
  const int h = H / 4, w = W / 4;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * h * w) return;
  const int x = i % w, y = (i / w) % h;
  const size_t k = i / ((size_t)w * h);
  const float* f = flow + (size_t)ids[k] * 2 * H * W;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float* q = f + (size_t)c * H * W + (size_t)(4 * y + 1) * W + 4 * x + 1;
    out[i * 2 + c] = (q[0] + q[1] + q[W] + q[W + 1]) * 0.25f * 0.25f;
  }
}

// prop_mask_in (:345-359): nearest 1/4 down-sampling (source pixel (4y, 4x)) of the input mask and of the frames' updated masks
// (channel 4 of the generator input) -> fp16 [n*h*w][8] with channels (m_in, m_updated)
__global__ void __launch_bounds__(256) pp_prop_masks_kernel(const __half* __restrict__ gen_in, int n, int H, int W, __half* __restrict__ out) {
  const int h = H / 4, w = W / 4;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * h * w) return;
  const int x = i % w, y = (i / w) % h;
  const size_t k = i / ((size_t)w * h);
  const __half* s = gen_in + ((k * H + (size_t)4 * y) * W + (size_t)4 * x) * 8;
  __align__(16) __half o[8] = {s[3], s[4], __half(), __half(), __half(), __half(), __half(), __half()};
  *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(o);
}

// condition tensor of the learnable feature propagation (:160-167): [cur (C), bilinear warp of prop by flow_prop (C), flow_prop (2), valid (1),
// masks of the current frame (2)] -> cond [P][pitch]; flows fp32 [P][2] at feature resolution; valid as in fbConsistencyCheck.  One thread =
// (pixel, 8 channels of the warp); the scalars ride on the first thread of a pixel.
__global__ void __launch_bounds__(256) pp_featprop_cond_kernel(const __half* __restrict__ prop, const __half* __restrict__ cur, int C, const float* __restrict__ fprop,
                                                               const float* __restrict__ fcheck, const __half* __restrict__ masks, int H, int W,
                                                               __half* __restrict__ cond, int pitch) {
  const int c8n = C >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)H * W * c8n) return;
  const int c8 = idx % c8n;
  const size_t p = idx / c8n;
  const int x = (int)(p % W), y = (int)(p / W);
  const float fx = fprop[p * 2], fy = fprop[p * 2 + 1];
  const float sx = (float)x + fx, sy = (float)y + fy;
  const float gx = floorf(sx), gy = floorf(sy);
  const int x0 = (int)gx, y0 = (int)gy;
  const float ax = sx - gx, ay = sy - gy;
  const float wgt[4] = {(1.f - ax) * (1.f - ay), ax * (1.f - ay), (1.f - ax) * ay, ax * ay};
  const int xs[4] = {x0, x0 + 1, x0, x0 + 1}, ys[4] = {y0, y0, y0 + 1, y0 + 1};
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (xs[k] < 0 || xs[k] >= W || ys[k] < 0 || ys[k] >= H) continue;
    const uint4 v = *reinterpret_cast<const uint4*>(prop + ((size_t)ys[k] * W + xs[k]) * C + c8 * 8);
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      acc[2 * j] += wgt[k] * f.x;
      acc[2 * j + 1] += wgt[k] * f.y;
    }
  }
  __half* o = cond + p * pitch;
  *reinterpret_cast<uint4*>(o + c8 * 8) = *reinterpret_cast<const uint4*>(cur + p * C + c8 * 8);
  __align__(16) __half2 wv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wv[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
  *reinterpret_cast<uint4*>(o + C + c8 * 8) = *reinterpret_cast<const uint4*>(wv);
  if (c8 == 0) {
    float b[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (xs[k] < 0 || xs[k] >= W || ys[k] < 0 || ys[k] >= H) continue;
      const float* q = fcheck + ((size_t)ys[k] * W + xs[k]) * 2;
      b[0] += wgt[k] * q[0];
      b[1] += wgt[k] * q[1];
    }
    const float dx = fx + b[0], dy = fy + b[1];
    const bool valid = dx * dx + dy * dy < 0.01f * (fx * fx + fy * fy + b[0] * b[0] + b[1] * b[1]) + 0.5f;
    o[2 * C] = __float2half_rn(fx);
    o[2 * C + 1] = __float2half_rn(fy);
    o[2 * C + 2] = __float2half_rn(valid ? 1.f : 0.f);
    o[2 * C + 3] = masks[p * 8];
    o[2 * C + 4] = masks[p * 8 + 1];
  }
}

// dst[p][coff .. coff+nch) = src[p][0 .. nch) for an 8-channel fp16 source (the 2 mask channels behind a feature concat)
__global__ void __launch_bounds__(256) pp_write_extra_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int pitch, int coff, int nch, size_t pixels) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  for (int c = 0; c < nch; ++c) dst[p * pitch + coff + c] = src[p * 8 + c];
}

// ---- P6 back half: soft split / composition, the sparse-window transformer (video/model/modules/sparse_transformer.py) -------------
// nn.Unfold(7, stride 3, padding 3) in TAP-MAJOR order: out[n][ty][tx][k*C + c] = in[n][3*ty - 3 + ky][3*tx - 3 + kx][c] (0 outside),
// k = ky*7 + kx.  (torch's order is c*49 + k; the linear layers that consume / produce it have their weights permuted on the host.)
__global__ void __launch_bounds__(256) pp_unfold7s3_kernel(const __half* __restrict__ in, int n, int h, int w, int C, int fh, int fw, __half* __restrict__ out,
                                                           int pitch, int gelu) {
  const int c8n = C >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * fh * fw * 49 * c8n) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int k = r % 49;
  r /= 49;
  const int tx = r % fw;
  r /= fw;
  const int ty = r % fh;
  const int img = r / fh;
  const int y = 3 * ty - 3 + k / 7, x = 3 * tx - 3 + k % 7;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (y >= 0 && y < h && x >= 0 && x < w) {
    v = *reinterpret_cast<const uint4*>(in + (((size_t)img * h + y) * w + x) * C + c8 * 8);
    if (gelu) {   // F.gelu (erf form) of FusionFeedForward.fc2[0], applied while the folded map is unfolded again
      __half2* h2 = reinterpret_cast<__half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h2[j]);
        f.x = 0.5f * f.x * (1.f + erff(f.x * 0.70710678f));
        f.y = 0.5f * f.y * (1.f + erff(f.y * 0.70710678f));
        h2[j] = __floats2half2_rn(f.x, f.y);
      }
    }
  }
  *reinterpret_cast<uint4*>(out + (((size_t)img * fh + ty) * fw + tx) * pitch + (size_t)k * C + c8 * 8) = v;
}

// F.fold(kernel 7, stride 3, padding 3) of tap-major tokens: out[n][y][x][c] = sum over the (token, tap) pairs that cover pixel (y, x) of
// tok[n][ty][tx][k*C + c]; normalise != 0 divides by the number of covering pairs (FusionFeedForward :96-108), else plain sum (SoftComp :61-66)
__global__ void __launch_bounds__(256) pp_fold7s3_kernel(const __half* __restrict__ tok, int n, int fh, int fw, int pitch, int C, int h, int w, int normalise,
                                                         __half* __restrict__ out) {
  const int c8n = C >> 3;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * h * w * c8n) return;
  const int c8 = idx % c8n;
  size_t r = idx / c8n;
  const int x = r % w;
  r /= w;
  const int y = r % h;
  const int img = r / h;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int cnt = 0;
  for (int ky = (y + 3) % 3; ky < 7; ky += 3) {
    const int ty = (y + 3 - ky) / 3;
    if (ty < 0 || ty >= fh) continue;
    for (int kx = (x + 3) % 3; kx < 7; kx += 3) {
      const int tx = (x + 3 - kx) / 3;
      if (tx < 0 || tx >= fw) continue;
      ++cnt;
      const uint4 v = *reinterpret_cast<const uint4*>(tok + (((size_t)img * fh + ty) * fw + tx) * pitch + (size_t)(ky * 7 + kx) * C + c8 * 8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
  }
  const float s = (normalise && cnt > 0) ? 1.f / (float)cnt : 1.f;
  __align__(16) __half2 o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) o[j] = __floats2half2_rn(acc[2 * j] * s, acc[2 * j + 1] * s);
  *reinterpret_cast<uint4*>(out + idx * 8) = *reinterpret_cast<const uint4*>(o);
}

// nn.LayerNorm(512) over the channel axis of tokens [P][C] (eps 1e-5); one warp per token
__global__ void __launch_bounds__(256) pp_layernorm_kernel(const __half* __restrict__ x, size_t tokens, int C, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, __half* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= tokens) return;
  const __half* src = x + t * C;
  float s = 0.f, q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = __half2float(src[c]);
    s += v;
    q += v * v;
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const float m = s / (float)C, r = rsqrtf(fmaxf(q / (float)C - m * m, 0.f) + 1e-5f);
  for (int c = lane; c < C; c += 32) out[t * C + c] = __float2half_rn((__half2float(src[c]) - m) * r * gamma[c] + beta[c]);
}

// SparseWindowAttention.pool_layer: depthwise Conv2d(C, C, 4, stride 4, groups C) with its own (trained) weights w[c][4][4], bias b[c]
__global__ void __launch_bounds__(256) pp_pool4_kernel(const __half* __restrict__ x, int n, int H, int W, int C, const float* __restrict__ wgt,
                                                       const float* __restrict__ bias, __half* __restrict__ out) {
  const int ph = H / 4, pw = W / 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * ph * pw * C) return;
  const int c = idx % C;
  size_t r = idx / C;
  const int px = r % pw;
  r /= pw;
  const int py = r % ph;
  const int img = r / ph;
  float acc = bias[c];
  for (int ky = 0; ky < 4; ++ky)
    for (int kx = 0; kx < 4; ++kx) acc += wgt[c * 16 + ky * 4 + kx] * __half2float(x[(((size_t)img * H + 4 * py + ky) * W + 4 * px + kx) * C + c]);
  out[idx] = __float2half_rn(acc);
}

// SparseWindowAttention core (:168-283) after the q/k/v projections, on padded token maps [T][Hn][Wn][C] (C = heads * 128).
// One block = (window, head, query frame): its wh*ww = 45 queries attend
//   masked window   : for every frame in t_ind: the window's own 45 tokens, the 148 valid tokens of the four rolled windows (valid_ind indexes the
//                     concatenation tl | tr | bl | br, each in window order) and all pooled tokens of that frame (kp / vp [T][ph][pw][C]);
//   unmasked window : the 45 tokens of the same window and the same frame only.
// torch.roll(k, (sy, sx)) puts token ((Y - sy) mod Hn, (X - sx) mod Wn) at (Y, X); the four shifts are (-eh,-ew), (-eh,+ew), (+eh,-ew), (+eh,+ew).
// Thread = (query, quarter of the 128 head dims); online softmax in fp32; scale 1/sqrt(128).  Keys and values of the head go through
// shared memory in tiles of PP_KT rows, loaded once per block with 16-byte accesses and converted to fp32 there (every row is consumed by
// all 45 queries: straight from global memory that is 45 x 64 two-byte loads per row and thread quarter).  Reads of a tile are warp
// broadcasts (the 8 queries of a warp read the same 4 segments), so the loop is bound by its 64 FMAs per key and thread.
#define PP_WH 5
#define PP_WW 9
#define PP_WT (PP_WH * PP_WW)
#define PP_KT 32
__global__ void __launch_bounds__(192) pp_window_attention_kernel(const __half* __restrict__ q, const __half* __restrict__ k, const __half* __restrict__ v,
                                                                  const __half* __restrict__ kp, const __half* __restrict__ vp, int T, int Hn, int Wn, int C,
                                                                  int ph, int pw, const int* __restrict__ valid_ind, int n_valid, const int* __restrict__ t_ind,
                                                                  int n_tind, const int* __restrict__ win_masked, __half* __restrict__ out) {
  __shared__ __align__(16) float ks[PP_KT][128];
  __shared__ __align__(16) float vs[PP_KT][128];
  const int nww = Wn / PP_WW;
  const int win = blockIdx.x, head = blockIdx.y, tq = blockIdx.z;
  const int wy0 = (win / nww) * PP_WH, wx0 = (win % nww) * PP_WW;
  const int qi = threadIdx.x >> 2, part = threadIdx.x & 3;        // query 0..47 (45 used), 32 dims each
  const bool active = qi < PP_WT;
  const int eh = (PP_WH + 1) / 2, ew = (PP_WW + 1) / 2;
  const size_t plane = (size_t)Hn * Wn;
  const int ch = head * 128;
  float qv[32], acc[32];
  float mx = -1e30f, den = 0.f;
  {   // idle threads (queries 45..47) run along with zero queries: they take part in the tile loads, the barriers and the shuffles, and never store
    const __half* src = q + (((size_t)tq * Hn + wy0 + (active ? qi / PP_WW : 0)) * Wn + wx0 + (active ? qi % PP_WW : 0)) * C + ch + part * 32;
#pragma unroll
    for (int d8 = 0; d8 < 4; ++d8) {
      const uint4 raw = *reinterpret_cast<const uint4*>(src + d8 * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        qv[d8 * 8 + 2 * j] = active ? f.x : 0.f;
        qv[d8 * 8 + 2 * j + 1] = active ? f.y : 0.f;
      }
    }
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = 0.f;
  }
  const bool masked = win_masked[win] != 0;
  const int n_pool = ph * pw;
  const int per_frame = masked ? PP_WT + n_valid + n_pool : PP_WT;
  const int n_frames = masked ? n_tind : 1;
  const float scale = 0.08838834764831845f;     // 1 / sqrt(128)
  for (int fi = 0; fi < n_frames; ++fi) {
    const int tk = masked ? t_ind[fi] : tq;
    for (int s0 = 0; s0 < per_frame; s0 += PP_KT) {
      const int nk = min(PP_KT, per_frame - s0);
      for (int e = threadIdx.x; e < nk * 16; e += blockDim.x) {      // one 8-channel group of one key / value row per step
        const int row = e >> 4, c8 = (e & 15) * 8, s = s0 + row;
        size_t off;
        const __half *kb, *vb;
        if (s < PP_WT + n_valid) {
          int wy, wx, sy = 0, sx = 0;
          if (s < PP_WT) { wy = s / PP_WW; wx = s % PP_WW; }
          else {
            const int id = valid_ind[s - PP_WT], rr = id / PP_WT, o = id % PP_WT;
            wy = o / PP_WW; wx = o % PP_WW;
            sy = (rr < 2) ? -eh : eh;
            sx = (rr & 1) ? ew : -ew;
          }
          const int Y = ((wy0 + wy - sy) % Hn + Hn) % Hn, X = ((wx0 + wx - sx) % Wn + Wn) % Wn;
          off = ((size_t)tk * plane + (size_t)Y * Wn + X) * C + ch + c8;
          kb = k; vb = v;
        } else {
          off = ((size_t)tk * n_pool + (s - PP_WT - n_valid)) * C + ch + c8;
          kb = kp; vb = vp;
        }
        const uint4 rk = *reinterpret_cast<const uint4*>(kb + off), rv = *reinterpret_cast<const uint4*>(vb + off);
        const __half2 *hk = reinterpret_cast<const __half2*>(&rk), *hv = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fk = __half22float2(hk[j]), fv = __half22float2(hv[j]);
          ks[row][c8 + 2 * j] = fk.x; ks[row][c8 + 2 * j + 1] = fk.y;
          vs[row][c8 + 2 * j] = fv.x; vs[row][c8 + 2 * j + 1] = fv.y;
        }
      }
      __syncthreads();
      for (int r = 0; r < nk; ++r) {
        const float* kr = &ks[r][part * 32];
        const float* vr = &vs[r][part * 32];
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < 32; ++d) dot += qv[d] * kr[d];
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot *= scale;
        const float nm = fmaxf(mx, dot), corr = __expf(mx - nm), pe = __expf(dot - nm);
        den = den * corr + pe;
#pragma unroll
        for (int d = 0; d < 32; ++d) acc[d] = acc[d] * corr + pe * vr[d];
        mx = nm;
      }
      __syncthreads();
    }
  }
  if (active) {
    __half* dst = out + (((size_t)tq * Hn + wy0 + qi / PP_WW) * Wn + wx0 + qi % PP_WW) * C + ch + part * 32;
    const float inv = 1.f / den;
#pragma unroll
    for (int d8 = 0; d8 < 4; ++d8) {
      __align__(16) __half o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(acc[d8 * 8 + j] * inv);
      *reinterpret_cast<uint4*>(dst + d8 * 8) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

// decoder output -> (tanh(x) + 1) / 2 * 255 truncated to u8, RGB [n][H][W][3] (propainter_inpaint.py:340-347 before the mask composite)
__global__ void __launch_bounds__(256) pp_pred_to_rgb8_kernel(const __half* __restrict__ x, int cp, size_t pixels, uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fmul_rn(__fdiv_rn(__fadd_rn(tanhf(__half2float(x[i * cp + c])), 1.f), 2.f), 255.f);
    out[i * 3 + c] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
  }
}

// combine_flow (:338-348): out[n] = pred[n'] * m + flow[n] * (1 - m), planar fp32; reverse: the network ran on the flipped sequence (n' = N-1-n)
__global__ void __launch_bounds__(256) pp_rfc_combine_kernel(const __half* __restrict__ pred, int pp, const float* __restrict__ flow, const uint8_t* __restrict__ mask,
                                                             int N, size_t plane, int reverse, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * plane) return;
  const size_t n = i / plane, p = i % plane;
  const __half* q = pred + ((size_t)(reverse ? N - 1 - n : n) * plane + p) * pp;
  const float m = mask[p] > 0 ? 1.f : 0.f;
  const float* f = flow + n * 2 * plane;
  float* o = out + n * 2 * plane;
  o[p] = __half2float(q[0]) * m + f[p] * (1.f - m);
  o[plane + p] = __half2float(q[1]) * m + f[plane + p] * (1.f - m);
}

}  // namespace vsr
