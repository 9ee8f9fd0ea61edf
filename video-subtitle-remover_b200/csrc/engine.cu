// vsr_b200 engine: C-ABI, weight packing, workspace, window schedule and kernel launches for the
// STTN-auto hot path.  Single translation unit: the .cuh files hold the
// sm_100a kernels, this file is the host runtime around them.
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cufft.h>

#include "../../include/vsr_b200.h"
#include "attention.cuh"
#include "conv_igemm.cuh"
#include "conv_halo.cuh"
#include "elementwise.cuh"
#include "host_index.h"
#include "rt_ops.cuh"
#include "pp_ops.cuh"

namespace vsr {

// ------------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define CK(expr)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (expr);                                                                                  \
    if (e_ != cudaSuccess)                                                                                    \
      throw Error(VSR_ERR_CUDA, std::string(#expr) + " -> " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                    std::to_string(__LINE__) + ")");                                          \
  } while (0)
#define REQUIRE(cond, msg) \
  do {                     \
    if (!(cond)) throw Error(VSR_ERR_ARG, std::string(msg) + " [" #cond "]"); \
  } while (0)

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return VSR_OK;
  } catch (const Error& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return VSR_ERR_STATE;
  }
}

static std::string device_error_report() {
  unsigned int h[4] = {0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(h, g_dev_error, sizeof(h)) != cudaSuccess) return "";
  if (!h[0]) return "";
  char buf[160];
  snprintf(buf, sizeof(buf), " [device watchdog: code 0x%x block %u bar 0x%x thread %u]", h[0], h[1], h[2], h[3]);
  return buf;
}

// ------------------------------------------------------------------------------------------------ device memory
// bumped on every (re)allocation: captured CUDA graphs bake pointers in.  Process-wide (any engine's reallocation makes every
// engine re-capture, which is only conservative) and atomic, so engines on different threads do not race on it; one engine
// itself is single-threaded (one Python thread per GPU).
static std::atomic<uint64_t> g_alloc_generation{0};

struct DevBuf {
  void* p = nullptr;
  size_t n = 0;
  void ensure(size_t bytes) {
    if (bytes <= n) return;
    ++g_alloc_generation;
    if (p) CK(cudaFree(p));
    p = nullptr;
    n = 0;
    const size_t want = (bytes + 255) & ~(size_t)255;
    CK(cudaMalloc(&p, want));
    CK(cudaMemset(p, 0, want));
    CK(cudaDeviceSynchronize());  // the memset runs on the legacy stream; engine streams are non-blocking
    n = want;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

template <class T>
static void upload(DevBuf& b, const std::vector<T>& v, cudaStream_t s = nullptr) {
  b.ensure(v.size() * sizeof(T) + 16);
  CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
}

// ------------------------------------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) throw Error(VSR_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// fp16 tensor, dims innermost-first, strides in BYTES for dims 1..rank-1, SWIZZLE_128B, zero OOB fill.
static CUtensorMap make_map_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  CUtensorMap m;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gs[i] = strides_bytes[i];
  }
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string s = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") rank " + std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) s += " " + std::to_string(dims[i]);
    s += " box";
    for (int i = 0; i < rank; ++i) s += " " + std::to_string(box[i]);
    throw Error(VSR_ERR_CUDA, s);
  }
  return m;
}

// ------------------------------------------------------------------------------------------------ launch helpers
struct Ctx {
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  bool conv_2cta = true;  // VSR_CONV_2CTA=0 falls back to the single-CTA 128x256 tile kernel (A/B switch)
  bool attn_2cta = true;  // VSR_ATTN_2CTA=0: single-CTA score / PV kernels
  bool conv_halo = false;  // haloed-tile kernel for k x k convs with 256-wide Cout tiles (conv_halo.cuh); VSR_CONV_HALO=0 switches it off
  bool conv_halo_narrow = false;  // VSR_CONV_HALO_NARROW=1: also the 64 / 128-wide Cout tiles on the haloed kernel (measured slower)
  int conv_cluster = 2;    // VSR_CONV_CLUSTER=4: two CTA pairs per cluster share the weights by TMA multicast (conv_halo.cuh)
  int conv_halo_base_off = 0;  // the descriptor's base-offset field stays 0: the tensor core swizzles on absolute smem address bits
                               // (measured, profiles/gpu_session_r2_s2_summary.txt: with the field set three conv cases fail)
  bool conv_tma_store = true;   // fp16-only conv outputs leave through a smem-staged TMA store (UTMASTG); VSR_CONV_TMA_STORE=0: per-thread stores
  bool direct_conv_smem = true;  // VSR_DIRECT_CONV_SMEM=0: the tiny-channel direct conv without shared-memory weights (A/B switch)
  bool conv_prefetch = false;  // VSR_CONV_PREFETCH=1: next-tile L2 prefetch in the conv producers (measured neutral)
  bool attn_lpt = true;   // VSR_ATTN_LPT=0: round-robin tile order in the score / PV launches
  bool attn_direct = true; // single-pass P = exp2(logit) for the heads without split-K (no S, no softmax kernel); VSR_ATTN_DIRECT=0: S + softmax kernel
  bool attn_fused = false; // VSR_ATTN_FUSED=1: two-pass score kernels without S (measured slower: the P pass is epilogue-bound)
  // In-situ profile (vsr_sttn_profile): with `prof` set, every ProfScope brackets its launches with a pair of events on the
  // stream; classes are the VSR_PROF_* ids of include/vsr_b200.h.
  bool prof = false;
  struct ProfRec { int cls; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
};

struct ProfScope {
  Ctx& c;
  Ctx::ProfRec r{0, nullptr, nullptr};
  ProfScope(Ctx& ctx, int cls) : c(ctx) {
    if (!c.prof) return;
    r.cls = cls;
    CK(cudaEventCreate(&r.a));
    CK(cudaEventCreate(&r.b));
    CK(cudaEventRecord(r.a, c.stream));
  }
  ~ProfScope() {
    if (!r.a) return;
    cudaEventRecord(r.b, c.stream);
    c.prof_recs.push_back(r);
  }
};

static bool env_flag(const char* name, bool dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) != 0 : dflt;
}

template <class P>
static void launch_tc(Ctx& c, const typename P::Params& prm, int ntiles) {
  static bool configured = false;
  constexpr int smem = tc_smem_bytes<P>();
  if (!configured) {
    CK(cudaFuncSetAttribute(tc_gemm_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (ntiles <= 0) return;
  const int grid = ntiles < c.sms ? ntiles : c.sms;
  tc_gemm_kernel<P><<<grid, TC_THREADS, smem, c.stream>>>(prm);
  CK(cudaGetLastError());
  ++c.launches;
}

template <class P>
static void launch_tc2(Ctx& c, const typename P::Params& prm, int ntiles) {
  static bool configured = false;
  constexpr int smem = tc2_smem_bytes<P>();
  if (!configured) {
    CK(cudaFuncSetAttribute(tc_gemm2_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  if (ntiles <= 0) return;
  const int pairs = c.sms / 2;
  const int grid = 2 * (ntiles < pairs ? ntiles : pairs);
  tc_gemm2_kernel<P><<<grid, tc2_threads<P>(), smem, c.stream>>>(prm);
  CK(cudaGetLastError());
  ++c.launches;
}

// ------------------------------------------------------------------------------------------------ conv layers
struct ConvLayer {
  DevBuf w, b;
  int cin = 0, cout = 0, cout_pad = 0, ntaps = 0, K = 0, bn = 0;
  int pitch = 0;  // channel pitch of the input tensor when it differs from cin (a channel-slice view); 0 = cin
  int8_t dy[81] = {0}, dx[81] = {0};
};

static int pad_cout(int cout) {
  if (cout <= 16) return 16;
  if (cout <= 64) return 64;
  if (cout <= 128) return 128;
  return (cout + 255) / 256 * 256;
}

// torch [Cout,Cin,k,k] fp32 -> fp16 [Cout_pad][tap][Cin], tap = ky*k+kx, offsets (ky-k/2)*dil.
static void pack_conv(ConvLayer& L, const float* w, const float* bias, int cout, int cin, int k, int dil, cudaStream_t s) {
  REQUIRE(cin % 64 == 0, "tcgen05 conv needs Cin multiple of 64");
  REQUIRE(k == 1 || k == 3, "kernel size 1 or 3");
  L.cin = cin; L.cout = cout; L.cout_pad = pad_cout(cout); L.ntaps = k * k; L.K = k * k * cin;
  L.bn = L.cout_pad < 256 ? L.cout_pad : 256;
  std::vector<__half> hw((size_t)L.cout_pad * L.K, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx)
          hw[(size_t)co * L.K + (ky * k + kx) * cin + ci] = __float2half_rn(w[(((size_t)co * cin + ci) * k + ky) * k + kx]);
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      L.dy[ky * k + kx] = (int8_t)((ky - k / 2) * dil);
      L.dx[ky * k + kx] = (int8_t)((kx - k / 2) * dil);
    }
  std::vector<float> hb(L.cout_pad, 0.f);
  for (int i = 0; i < cout; ++i) hb[i] = bias ? bias[i] : 0.f;
  upload(L.w, hw, s);
  upload(L.b, hb, s);
}

// conv3x3 stride 2 pad 1 over [H,W,Cin] == conv 2x2 (taps dy,dx in {-1,0}) stride 1 over the
// space-to-depth tensor [H/2,W/2,4*Cin] (channel = (py*2+px)*Cin + ci): input row 2*yo+ky-1 =
// 2*(yo+dyb)+py  =>  ky = 2*dyb + py + 1.  Unused (tap, parity) combinations get zero weights.
static void pack_conv_s2d(ConvLayer& L, const float* w, const float* bias, int cout, int cin, cudaStream_t s) {
  REQUIRE((4 * cin) % 64 == 0, "s2d conv needs 4*Cin multiple of 64");
  L.cin = 4 * cin; L.cout = cout; L.cout_pad = pad_cout(cout); L.ntaps = 4; L.K = 16 * cin;
  L.bn = L.cout_pad < 256 ? L.cout_pad : 256;
  std::vector<__half> hw((size_t)L.cout_pad * L.K, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int dyb = -1; dyb <= 0; ++dyb)
      for (int dxb = -1; dxb <= 0; ++dxb)
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px) {
            const int ky = 2 * dyb + py + 1, kx = 2 * dxb + px + 1;
            if (ky < 0 || ky > 2 || kx < 0 || kx > 2) continue;
            const int tap = (dyb + 1) * 2 + (dxb + 1);
            for (int ci = 0; ci < cin; ++ci)
              hw[(size_t)co * L.K + (size_t)tap * 4 * cin + (py * 2 + px) * cin + ci] =
                  __float2half_rn(w[(((size_t)co * cin + ci) * 3 + ky) * 3 + kx]);
          }
  for (int dyb = -1; dyb <= 0; ++dyb)
    for (int dxb = -1; dxb <= 0; ++dxb) {
      L.dy[(dyb + 1) * 2 + (dxb + 1)] = (int8_t)dyb;
      L.dx[(dyb + 1) * 2 + (dxb + 1)] = (int8_t)dxb;
    }
  std::vector<float> hb(L.cout_pad, 0.f);
  for (int i = 0; i < cout; ++i) hb[i] = bias ? bias[i] : 0.f;
  upload(L.w, hw, s);
  upload(L.b, hb, s);
}

// General stride-1 dense conv for the graph runtime: kernel kh x kw, dilation, explicit top/left padding; the
// input tensor has `cin_pitch` (multiple of 64) channels of which the first `cin` are real.
static void pack_conv_general(ConvLayer& L, const float* w, const float* bias, int cout, int cin, int cin_pitch, int kh, int kw,
                              int dil, int pad_t, int pad_l, cudaStream_t s, bool split = false) {
  REQUIRE(cin_pitch % 8 == 0 && cin <= cin_pitch, "input channel pitch must be a multiple of 8 (16-byte TMA strides)");
  const int taps = kh * kw;
  REQUIRE((split ? 2 : 1) * taps <= 81, "kernels up to 81 taps (40 with split weights)");
  const int cin_k = (cin + 63) / 64 * 64;  // K extent per tap: whole 64-channel TMA boxes (weights beyond cin are zero)
  REQUIRE(cin_k <= cin_pitch || cin_pitch % 64 == 0, "a channel slice must end on the tensor's 64-channel grid");
  // split: every tap appears twice with the same offset, once with hi = fp16(w) and once with lo = fp16(w - hi): the tensor cores then see
  // the weight to ~22 bits while the activations stay fp16 (the detector's head is limited by weight rounding)
  L.cin = cin_k; L.pitch = cin_pitch; L.cout = cout; L.cout_pad = pad_cout(cout); L.ntaps = (split ? 2 : 1) * taps; L.K = L.ntaps * cin_k;
  L.bn = L.cout_pad < 256 ? L.cout_pad : 256;
  std::vector<__half> hw((size_t)L.cout_pad * L.K, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int ky = 0; ky < kh; ++ky)
        for (int kx = 0; kx < kw; ++kx) {
          const float v = w[(((size_t)co * cin + ci) * kh + ky) * kw + kx];
          const __half hi = __float2half_rn(v);
          hw[(size_t)co * L.K + (size_t)(ky * kw + kx) * cin_k + ci] = hi;
          if (split) hw[(size_t)co * L.K + (size_t)(taps + ky * kw + kx) * cin_k + ci] = __float2half_rn(v - __half2float(hi));
        }
  for (int rep = 0; rep < (split ? 2 : 1); ++rep)
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx) {
        L.dy[rep * taps + ky * kw + kx] = (int8_t)(ky * dil - pad_t);
        L.dx[rep * taps + ky * kw + kx] = (int8_t)(kx * dil - pad_l);
      }
  std::vector<float> hb(L.cout_pad, 0.f);
  for (int i = 0; i < cout; ++i) hb[i] = bias ? bias[i] : 0.f;
  upload(L.w, hw, s);
  upload(L.b, hb, s);
}

// conv3x3 stride 2 pad 1 over a tensor with channel pitch cin_pitch, as 2x2 taps over its space-to-depth form
static void pack_conv_s2d_pitch(ConvLayer& L, const float* w, const float* bias, int cout, int cin, int cin_pitch, cudaStream_t s) {
  REQUIRE((4 * cin_pitch) % 64 == 0, "s2d conv needs 4*pitch multiple of 64");
  L.cin = 4 * cin_pitch; L.cout = cout; L.cout_pad = pad_cout(cout); L.ntaps = 4; L.K = 16 * cin_pitch;
  L.bn = L.cout_pad < 256 ? L.cout_pad : 256;
  std::vector<__half> hw((size_t)L.cout_pad * L.K, __float2half(0.f));
  for (int co = 0; co < cout; ++co)
    for (int dyb = -1; dyb <= 0; ++dyb)
      for (int dxb = -1; dxb <= 0; ++dxb)
        for (int py = 0; py < 2; ++py)
          for (int px = 0; px < 2; ++px) {
            const int ky = 2 * dyb + py + 1, kx = 2 * dxb + px + 1;
            if (ky < 0 || ky > 2 || kx < 0 || kx > 2) continue;
            const int tap = (dyb + 1) * 2 + (dxb + 1);
            for (int ci = 0; ci < cin; ++ci)
              hw[(size_t)co * L.K + (size_t)tap * 4 * cin_pitch + (py * 2 + px) * cin_pitch + ci] =
                  __float2half_rn(w[(((size_t)co * cin + ci) * 3 + ky) * 3 + kx]);
          }
  for (int dyb = -1; dyb <= 0; ++dyb)
    for (int dxb = -1; dxb <= 0; ++dxb) {
      L.dy[(dyb + 1) * 2 + (dxb + 1)] = (int8_t)dyb;
      L.dx[(dyb + 1) * 2 + (dxb + 1)] = (int8_t)dxb;
    }
  std::vector<float> hb(L.cout_pad, 0.f);
  for (int i = 0; i < cout; ++i) hb[i] = bias ? bias[i] : 0.f;
  upload(L.w, hw, s);
  upload(L.b, hb, s);
}

static void choose_tile(int H, int W, int& tw, int& th) {
  double best = -1;
  tw = 32; th = 4;
  for (int w = 1; w <= 128 && w <= 256; ++w) {
    if (w > W && w != 1) break;
    const int h = 128 / w;
    if (h < 1) break;
    const int hh = h > H ? H : h;
    const long long covered = (long long)((W + w - 1) / w) * ((H + hh - 1) / hh);
    const double util = (double)W * H / (double)(covered * 128);
    if (util >= best - 1e-9) { best = util; tw = w; th = hh; }  // ties: prefer the widest tile
  }
}

struct ConvIO {
  const __half* in = nullptr;
  int T = 0, H = 0, W = 0;
  int flags = 0;
  __half* out16 = nullptr;
  int out16_pitch = 0, out16_coff = 0;
  float* out32 = nullptr;
  const float* res32 = nullptr;
  float* comps = nullptr;
  const int* frame_idx = nullptr;
  const int* first_visit = nullptr;
  const uint8_t* det_mask = nullptr;
  const uchar4* det_rgb = nullptr;
  float alpha = 1.f, bias_scale = 1.f;  // CONV_SCALED (graph runtime)
  int* overflow = nullptr;
  int crop_t = 0, crop_l = 0, out_H = 0, out_W = 0;  // out_H == 0: the output grid is the input grid
};

static void run_conv(Ctx& c, const ConvLayer& L, const ConvIO& io) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  int tw, th;
  choose_tile(io.H, io.W, tw, th);
  {
    const uint64_t pitch = L.pitch ? L.pitch : L.cin;
    const uint64_t dims[4] = {(uint64_t)L.cin, (uint64_t)io.W, (uint64_t)io.H, (uint64_t)io.T};
    const uint64_t str[3] = {pitch * 2, (uint64_t)io.W * pitch * 2, (uint64_t)io.H * io.W * pitch * 2};
    const uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
    p.in_map = make_map_f16(io.in, 4, dims, str, box);
  }
  {
    const uint64_t dims[2] = {(uint64_t)L.K, (uint64_t)L.cout_pad};
    const uint64_t str[1] = {(uint64_t)L.K * 2};
    const uint32_t box[2] = {64, (uint32_t)L.bn};
    p.w_map = make_map_f16(L.w.p, 2, dims, str, box);
  }
  p.T = io.T; p.H = io.H; p.W = io.W;
  p.tile_w = tw; p.tile_h = th;
  p.tiles_x = (io.W + tw - 1) / tw;
  p.tiles_y = (io.H + th - 1) / th;
  p.n_tiles = L.cout_pad / L.bn;
  p.ntaps = L.ntaps;
  p.cin_chunks = L.cin / 64;
  p.cout = L.cout;
  p.flags = io.flags;
  p.prefetch = c.conv_prefetch ? 1 : 0;
  memcpy(p.tap_dy, L.dy, 81);
  memcpy(p.tap_dx, L.dx, 81);
  p.bias = L.b.as<float>();
  p.out16 = io.out16;
  p.out16_pitch = io.out16_pitch ? io.out16_pitch : L.cout;
  p.out16_coff = io.out16_coff;
  p.out32 = io.out32;
  p.res32 = io.res32;
  p.comps = io.comps;
  p.frame_idx = io.frame_idx;
  p.first_visit = io.first_visit;
  p.det_mask = io.det_mask;
  p.det_rgb = io.det_rgb;
  p.alpha = io.alpha;
  p.bias_scale = io.bias_scale;
  p.overflow = io.overflow;
  p.crop_t = io.crop_t;
  p.crop_l = io.crop_l;
  p.out_H = io.out_H ? io.out_H : io.H;
  p.out_W = io.out_W ? io.out_W : io.W;
  if (io.flags & CONV_S2D_STORE) {
    REQUIRE(io.H % 2 == 0 && io.W % 2 == 0, "s2d store needs even H, W");
    p.out16_pitch = 4 * L.cout;
  }
  if (!(io.flags & CONV_FINAL)) REQUIRE(L.cout % 8 == 0, "Cout must be a multiple of 8");
  if (io.out32) REQUIRE(L.cout == L.cout_pad, "fp32 stream needs Cout == padded Cout");
  // fp16-only outputs on an un-cropped grid: TMA store epilogue (conv_store_tma); the map is finished below, once the tile shape is known
  const bool tma_store = c.conv_tma_store && io.out16 && !io.out32 && !(io.flags & (CONV_FINAL | CONV_S2D_STORE | CONV_SCALED)) && io.out_H == 0 &&
                         io.crop_t == 0 && io.crop_l == 0 && L.bn == 256 && c.conv_2cta && (p.out16_pitch % 8) == 0 && (p.out16_coff % 8) == 0;
  auto make_out_map = [&](int tw, int th) {
    const uint64_t od[4] = {(uint64_t)L.cout, (uint64_t)io.W, (uint64_t)io.H, (uint64_t)io.T};
    const uint64_t os[3] = {(uint64_t)p.out16_pitch * 2, (uint64_t)io.W * p.out16_pitch * 2, (uint64_t)io.H * io.W * p.out16_pitch * 2};
    const uint32_t ob[4] = {32, (uint32_t)tw, (uint32_t)th, 1};
    p.out_map = make_map_f16(io.out16 + p.out16_coff, 4, od, os, ob, CU_TENSOR_MAP_SWIZZLE_64B);
    p.flags |= CONV_TMA_STORE;
  };
  const int ntiles = p.T * p.tiles_y * p.tiles_x * p.n_tiles;
  // 64- and 128-wide Cout tiles (decoder convs) measured SLOWER on the haloed pair kernel than on the per-tap single-CTA one (decoder 6.1 vs
  // 5.2 ms per chunk, profiles/gpu_session_r2_s4c_summary.txt): kept behind VSR_CONV_HALO_NARROW=1
  if ((L.bn == 256 || (c.conv_halo_narrow && (L.bn == 128 || L.bn == 64))) && c.conv_2cta && c.conv_halo && !(io.flags & (CONV_FINAL | CONV_S2D_STORE)) &&
      L.ntaps > 1 && io.W >= 8) {
    int halo = 0;
    for (int i = 0; i < L.ntaps; ++i) halo = std::max(halo, std::max(std::abs((int)L.dy[i]), std::abs((int)L.dx[i])));
    if (halo <= 4) {
      // tiles of 8 x 16 pixels; tiles advance by the fewest rows that cover H in ceil(H / 16) steps (30 rows -> 2 x 15)
      const int tiles_y = (io.H + 15) / 16, valid_h = (io.H + tiles_y - 1) / tiles_y;
      const uint64_t pitch = L.pitch ? L.pitch : L.cin;
      const uint64_t dims[4] = {(uint64_t)L.cin, (uint64_t)io.W, (uint64_t)io.H, (uint64_t)io.T};
      const uint64_t str[3] = {pitch * 2, (uint64_t)io.W * pitch * 2, (uint64_t)io.H * io.W * pitch * 2};
      const uint32_t box[4] = {64, 16, (uint32_t)(16 + 2 * halo), 1};
      p.in_map = make_map_f16(io.in, 4, dims, str, box);
      const uint64_t wd[2] = {(uint64_t)L.K, (uint64_t)L.cout_pad};
      const uint64_t ws[1] = {(uint64_t)L.K * 2};
      const uint32_t wb[2] = {64, (uint32_t)(L.bn / 2)};
      p.w_map_half = make_map_f16(L.w.p, 2, wd, ws, wb);
      p.tile_w = 8; p.tile_h = valid_h;
      p.tiles_x = (io.W + 7) / 8;
      p.tiles_y = tiles_y;
      p.halo = halo;
      p.halo_base_off = c.conv_halo_base_off;
      p.scr_stride = 1024;
      if (tma_store && L.bn == 256) make_out_map(8, valid_h);
      p.halo_nb = halo_b_stages(halo, L.bn);
      const int smem = halo_smem_bytes(halo, L.bn);
      static int configured_smem[4] = {0, 0, 0, 0}, max_clusters4 = 0;
      auto configure = [&](int slot, const void* fn) {
        if (smem > configured_smem[slot]) {
          CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
          configured_smem[slot] = smem;
        }
      };
      const int pair_tiles_m = (p.T * p.tiles_y * p.tiles_x + 1) / 2;
      const int groups = pair_tiles_m * p.n_tiles;
      const int grid = 2 * std::min(groups, c.sms / 2);
      if (L.bn == 256 && c.conv_cluster == 4) {
        configure(3, (const void*)conv_halo_kernel<4, 256>);
        if (max_clusters4 == 0) {
          // how many 4-CTA clusters the device holds at once (GPCs with an odd number of TPCs leave SMs without a cluster)
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(4 * (c.sms / 4)); cfg.blockDim = dim3(HALO_THREADS); cfg.dynamicSmemBytes = (size_t)smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          if (cudaOccupancyMaxActiveClusters(&max_clusters4, conv_halo_kernel<4, 256>, &cfg) != cudaSuccess || max_clusters4 <= 0) {
            cudaGetLastError();
            max_clusters4 = c.sms / 4;
          }
          if (getenv("VSR_DEBUG_CLUSTERS")) fprintf(stderr, "[vsr] conv_halo: %d clusters of 4 fit on %d SMs\n", max_clusters4, c.sms);
        }
        const uint32_t qb[2] = {64, 64};
        p.w_map_quarter = make_map_f16(L.w.p, 2, wd, ws, qb);
        const int groups4 = ((pair_tiles_m + 1) / 2) * p.n_tiles;
        conv_halo_kernel<4, 256><<<4 * std::min(groups4, max_clusters4), HALO_THREADS, smem, c.stream>>>(p);
      } else if (L.bn == 256) {
        configure(0, (const void*)conv_halo_kernel<2, 256>);
        conv_halo_kernel<2, 256><<<grid, HALO_THREADS, smem, c.stream>>>(p);
      } else if (L.bn == 128) {
        configure(1, (const void*)conv_halo_kernel<2, 128>);
        conv_halo_kernel<2, 128><<<grid, HALO_THREADS, smem, c.stream>>>(p);
      } else {
        configure(2, (const void*)conv_halo_kernel<2, 64>);
        conv_halo_kernel<2, 64><<<grid, HALO_THREADS, smem, c.stream>>>(p);
      }
      CK(cudaGetLastError());
      ++c.launches;
      return;
    }
  }
  if (L.bn == 256 && c.conv_2cta && !(io.flags & CONV_FINAL)) {
    const uint64_t dims[2] = {(uint64_t)L.K, (uint64_t)L.cout_pad};
    const uint64_t str[1] = {(uint64_t)L.K * 2};
    const uint32_t box[2] = {64, 128};
    p.w_map_half = make_map_f16(L.w.p, 2, dims, str, box);
    p.scr_stride = 32 * 33;
    if (tma_store) make_out_map(tw, th);
    launch_tc2<Conv2Policy>(c, p, ((p.T * p.tiles_y * p.tiles_x + 1) / 2) * p.n_tiles);
    return;
  }
  switch (L.bn) {
    case 256: launch_tc<ConvPolicy<256>>(c, p, ntiles); break;
    case 128: launch_tc<ConvPolicy<128>>(c, p, ntiles); break;
    case 64: launch_tc<ConvPolicy<64>>(c, p, ntiles); break;
    case 16: launch_tc<ConvPolicy<16>>(c, p, ntiles); break;
    default: throw Error(VSR_ERR_STATE, "unsupported BN");
  }
}

// ------------------------------------------------------------------------------------------------ attention
struct TileSchedule {
  DevBuf dev;
  int stride = 0;
};
struct AttnWorkspace {
  DevBuf S[ATTN_MAX_HEADS], P[ATTN_MAX_HEADS], rowsum[ATTN_MAX_HEADS], rowmax_part[ATTN_MAX_HEADS], rowsum_part[ATTN_MAX_HEADS];
  DevBuf overflow;   // [0]: a direct head met a key that beats the token's own logit by more than fp16 holds
  DevBuf rowshift[ATTN_MAX_HEADS];
  std::map<std::string, std::unique_ptr<TileSchedule>> schedules;  // LPT tile orders, keyed by the problem shapes
};

// Longest-processing-time-first assignment of tiles (cost in K-chunks + a fixed per-tile overhead) to `bins`
// CTAs / CTA pairs.  The attention launches mix tiles of 1 to >100 chunks; round robin left the busiest CTA
// ~28 % above the mean.  Cached per shape signature so the device pointer is stable under CUDA graphs.
static const TileSchedule& lpt_schedule(AttnWorkspace& ws, const std::string& key, const std::vector<int>& cost, int bins,
                                        cudaStream_t stream) {
  auto it = ws.schedules.find(key);
  if (it != ws.schedules.end()) return *it->second;
  const int n = (int)cost.size();
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  std::vector<std::vector<int>> lists(bins);
  typedef std::pair<long long, int> LB;  // (load, bin)
  std::priority_queue<LB, std::vector<LB>, std::greater<LB>> pq;
  for (int b = 0; b < bins; ++b) pq.push({0, b});
  for (int i : idx) {
    LB lb = pq.top();
    pq.pop();
    lists[lb.second].push_back(i);
    pq.push({lb.first + cost[i], lb.second});
  }
  size_t maxlen = 1;
  for (auto& l : lists) maxlen = std::max(maxlen, l.size() + 1);
  std::vector<int> flat((size_t)bins * maxlen, -1);
  for (int b = 0; b < bins; ++b)
    for (size_t j = 0; j < lists[b].size(); ++j) flat[(size_t)b * maxlen + j] = lists[b][j];
  auto sch = std::make_unique<TileSchedule>();
  sch->stride = (int)maxlen;
  upload(sch->dev, flat, stream);
  return *(ws.schedules[key] = std::move(sch));
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct AttnSegment {  // one attention problem: frames [first, first + T) of the buffers (a window of the schedule)
  int first, T;
};

// qkv: NHWC fp16 [frames,H,W,pitch]; q/k/v channel offsets q_off/k_off/v_off; head i uses +i*dk.  Every
// (segment, patch geometry) pair is an independent attention problem; all of them run in the same three
// launches (scores, softmax, PV).
static void run_attention(Ctx& c, AttnWorkspace& ws, const __half* qkv, int pitch, int q_off, int k_off, int v_off,
                          const std::vector<AttnSegment>& segs, int H, int W, int C, int n_patch, const int* pw, const int* ph,
                          __half* out, int out_pitch) {
  REQUIRE(n_patch >= 1 && n_patch <= 4, "1..4 heads");
  const int dk = C / n_patch;
  REQUIRE(dk == 64, "head width must be 64 channels (one SWIZZLE_128B row)");
  const int nent = (int)segs.size() * n_patch;
  REQUIRE(nent >= 1 && nent <= ATTN_MAX_HEADS, "too many attention problems in one launch");
  ScoreParams sp;
  PVParams pp;
  memset(&sp, 0, sizeof(sp));
  memset(&pp, 0, sizeof(pp));
  // order the problems by descending work so the long tiles are scheduled first
  struct Ent { int seg, patch; double work; };
  std::vector<Ent> ents;
  for (int g = 0; g < (int)segs.size(); ++g)
    for (int i = 0; i < n_patch; ++i) {
      const double n = (double)segs[g].T * (H / ph[i]) * (W / pw[i]);
      ents.push_back({g, i, n * n * 64.0 * pw[i] * ph[i]});
    }
  std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.work > b.work; });
  // split-K only where one tile would be a large share of a CTA's fair load (in 64-wide K-chunk units)
  double total_units = 0;
  for (const Ent& e : ents) {
    const int i = e.patch;
    const int ntok_p = segs[e.seg].T * (H / ph[i]) * next_pow2(W / pw[i]);
    const int ntt = (ntok_p + 127) / 128;
    total_units += (double)ntt * ((ntt + 1) / 2) * pw[i] * ph[i];
  }
  const int split_target = std::max(8, (int)(total_units / c.sms / 4.0));
  int score_work = 0, pv_work = 0, score_work2 = 0, pv_work2 = 0, max_rows = 0, workB = 0, workB2 = 0;
  bool any_unfused = false;
  for (int s = 0; s < nent; ++s) {
    const int i = ents[s].patch;
    const AttnSegment& sg = segs[ents[s].seg];
    const size_t frame_elems = (size_t)H * W * pitch;
    AttnHead h;
    memset(&h, 0, sizeof(h));
    REQUIRE(W % pw[i] == 0 && H % ph[i] == 0, "patch must divide the feature map");
    h.pw = pw[i]; h.ph = ph[i];
    h.ow = W / pw[i]; h.oh = H / ph[i];
    h.owp = next_pow2(h.ow);
    REQUIRE(h.owp <= 64, "more than 64 patches per row is not supported");
    h.npos = pw[i] * ph[i];
    h.toh_total = sg.T * h.oh;
    const int ntok_p = h.toh_total * h.owp;
    h.ntt = (ntok_p + 127) / 128;
    h.nk64 = (ntok_p + 63) / 64;
    const int tiles = h.ntt * ((h.ntt + 1) / 2);  // 128 queries x 256 keys per CTA tile
    const int splits = h.npos > split_target ? (h.npos + split_target - 1) / split_target : 1;
    h.chunks_per_split = (h.npos + splits - 1) / splits;
    h.splits = (h.npos + h.chunks_per_split - 1) / h.chunks_per_split;
    h.fused = (h.splits == 1 && c.attn_direct && c.attn_2cta) ? 2 : (h.splits == 1 && c.attn_fused) ? 1 : 0;
    h.npairs = (h.ntt + 1) / 2;
    any_unfused |= !h.fused;
    h.score_work_begin = score_work;
    h.score_work_begin2 = score_work2;
    if (h.fused != 2) {   // direct heads have no pass A
      score_work += tiles * h.splits;
      score_work2 += ((h.ntt + 1) / 2) * ((h.ntt + 1) / 2) * h.splits;
    }
    h.pv_ntiles = (h.npos + 3) / 4;
    h.pv_work_begin = pv_work;
    pv_work += h.ntt * h.pv_ntiles;
    h.pv_work_begin2 = pv_work2;
    pv_work2 += ((h.ntt + 1) / 2) * h.pv_ntiles;
    h.scoreB_begin = workB;
    h.scoreB_begin2 = workB2;
    if (h.fused) {
      workB += tiles;
      workB2 += ((h.ntt + 1) / 2) * ((h.ntt + 1) / 2);
    }
    h.ldS = h.ntt * 128;
    h.ldP = h.ntt * 128;
    h.scale_log2e = (float)(1.4426950408889634 / std::sqrt((double)dk * h.npos));
    const size_t rows = (size_t)h.ntt * 128;
    h.slabS = (long long)rows * h.ldS;
    if (!h.fused) ws.S[s].ensure((size_t)h.splits * rows * h.ldS * sizeof(float));
    ws.P[s].ensure(rows * h.ldP * sizeof(__half));
    ws.rowsum[s].ensure(rows * sizeof(float));
    ws.rowmax_part[s].ensure(rows * h.npairs * sizeof(float));
    ws.rowsum_part[s].ensure(rows * h.npairs * sizeof(float));
    h.S = h.fused ? nullptr : ws.S[s].as<float>();
    h.P = ws.P[s].as<__half>();
    h.rowsum = ws.rowsum[s].as<float>();
    h.rowmax_part = ws.rowmax_part[s].as<float>();
    ws.rowshift[s].ensure(rows * sizeof(float));
    h.rowshift = ws.rowshift[s].as<float>();
    h.rowsum_part = ws.rowsum_part[s].as<float>();
    if ((int)rows > max_rows) max_rows = (int)rows;
    // 5-D views {64 ch, px, ow, py, toh} of this segment's frames
    const __half* base = qkv + (size_t)sg.first * frame_elems;
    const uint64_t dims[5] = {64, (uint64_t)h.pw, (uint64_t)h.ow, (uint64_t)h.ph, (uint64_t)h.toh_total};
    const uint64_t str[4] = {(uint64_t)pitch * 2, (uint64_t)h.pw * pitch * 2, (uint64_t)W * pitch * 2,
                             (uint64_t)h.ph * W * pitch * 2};
    const uint32_t box_qk[5] = {64, 1, (uint32_t)h.owp, 1, (uint32_t)(128 / h.owp)};
    const uint32_t box_v[5] = {64, 1, (uint32_t)h.owp, 1, (uint32_t)(64 / h.owp)};
    sp.qmap[s] = make_map_f16(base + q_off + i * dk, 5, dims, str, box_qk);
    sp.kmap[s] = make_map_f16(base + k_off + i * dk, 5, dims, str, box_qk);
    pp.vmap[s] = make_map_f16(base + v_off + i * dk, 5, dims, str, box_v);
    const uint64_t pd[2] = {(uint64_t)h.ldP, (uint64_t)rows};
    const uint64_t ps[1] = {(uint64_t)h.ldP * 2};
    const uint32_t pb[2] = {64, 128};
    pp.pmap[s] = make_map_f16(h.P, 2, pd, ps, pb);
    sp.h[s] = h;
    pp.h[s] = h;
    pp.coff[s] = i * dk;
    pp.out_off[s] = (long long)sg.first * H * W * out_pitch;
  }
  sp.nheads = pp.nheads = nent;
  sp.total_work = score_work;
  pp.total_work = pv_work;
  sp.total_work2 = score_work2;
  pp.total_work2 = pv_work2;
  pp.T = 0; pp.H = H; pp.W = W;
  pp.out = out;
  pp.out_pitch = out_pitch;
  if (c.attn_lpt) {
    const bool two = c.attn_2cta;
    const int bins = two ? c.sms / 2 : c.sms;
    std::string key = two ? "2" : "1";
    std::vector<int> cs, cp;
    for (int s2 = 0; s2 < nent; ++s2) {
      const AttnHead& h = sp.h[s2];
      key += "|" + std::to_string(h.ntt) + "," + std::to_string(h.npos) + "," + std::to_string(h.splits) + "," + std::to_string(h.nk64);
      key += h.fused == 2 ? "d" : "";
      const int nq = two ? (h.ntt + 1) / 2 : h.ntt, nk2 = (h.ntt + 1) / 2;
      for (int q = 0; q < nq && h.fused != 2; ++q)
        for (int kj = 0; kj < nk2; ++kj)
          for (int spl = 0; spl < h.splits; ++spl) cs.push_back(std::min(h.chunks_per_split, h.npos - spl * h.chunks_per_split) + 8);
      for (int m = 0; m < nq; ++m)
        for (int ni = 0; ni < h.pv_ntiles; ++ni) cp.push_back(h.nk64 + 3);
    }
    if (!cs.empty()) {
      const TileSchedule& ss = lpt_schedule(ws, "S" + key, cs, std::min(bins, (int)cs.size()), c.stream);
      sp.order = ss.dev.as<int>();
      sp.order_stride = ss.stride;
    }
    const TileSchedule& ps = lpt_schedule(ws, "P" + key, cp, std::min(bins, (int)cp.size()), c.stream);
    pp.order = ps.dev.as<int>();
    pp.order_stride = ps.stride;
  }
  sp.totalB = workB;
  sp.totalB2 = workB2;
  ws.overflow.ensure(16);
  sp.overflow = ws.overflow.as<int>();
  pp.overflow = ws.overflow.as<int>();
  if (workB > 0 && c.attn_direct) {   // row shifts of the direct heads: each token's logit with itself
    DiagParams dp;
    memset(&dp, 0, sizeof(dp));
    int rows_total = 0;
    for (int s2 = 0; s2 < nent; ++s2) {
      dp.row_begin[s2] = rows_total;
      dp.h[s2] = sp.h[s2];
      const AttnSegment& sg = segs[ents[s2].seg];
      const __half* base = qkv + (size_t)sg.first * H * W * pitch;
      dp.q[s2] = base + q_off + ents[s2].patch * dk;
      dp.k[s2] = base + k_off + ents[s2].patch * dk;
      if (sp.h[s2].fused == 2) rows_total += sp.h[s2].ntt * 128;
    }
    for (int s2 = nent; s2 <= ATTN_MAX_HEADS; ++s2) dp.row_begin[s2] = rows_total;
    dp.nheads = nent; dp.pitch = pitch; dp.W = W;
    if (rows_total > 0) {
      ProfScope ps_(c, VSR_PROF_SOFTMAX);
      attn_diag_kernel<<<(rows_total * 32 + 255) / 256, 256, 0, c.stream>>>(dp);
      CK(cudaGetLastError());
      ++c.launches;
    }
  }
  sp.pass = 0;  // pass A: per-tile row maxima (fused problems) / fp32 S slabs (split-K problems)
  if ((c.attn_2cta ? score_work2 : score_work) > 0) {
    ProfScope ps_(c, VSR_PROF_SCORE);
    if (c.attn_2cta) launch_tc2<Score2Policy>(c, sp, score_work2);
    else launch_tc<ScorePolicy>(c, sp, score_work);
  }
  if (any_unfused) {
    ProfScope ps_(c, VSR_PROF_SOFTMAX);
    int max_cols = 0, rows_total = 0;
    for (int s2 = 0; s2 < nent; ++s2) {
      sp.softmax_row_begin[s2] = rows_total;
      if (!sp.h[s2].fused) {
        max_cols = std::max(max_cols, sp.h[s2].nk64 * 64);
        rows_total += sp.h[s2].ntt * 128;
      }
    }
    for (int s2 = nent; s2 <= ATTN_MAX_HEADS; ++s2) sp.softmax_row_begin[s2] = rows_total;
    const dim3 grid(rows_total);
    if (max_cols <= 1024) softmax_rows_kernel<1><<<grid, 256, 0, c.stream>>>(sp);
    else if (max_cols <= 2048) softmax_rows_kernel<2><<<grid, 256, 0, c.stream>>>(sp);
    else if (max_cols <= 5120) softmax_rows_kernel<5><<<grid, 256, 0, c.stream>>>(sp);
    else if (max_cols <= 10240) softmax_rows_kernel<10><<<grid, 256, 0, c.stream>>>(sp);
    else if (max_cols <= 20480) softmax_rows_kernel<20><<<grid, 256, 0, c.stream>>>(sp);
    else throw Error(VSR_ERR_ARG, "attention rows longer than 20480 tokens are not supported");
    CK(cudaGetLastError());
    ++c.launches;
  }
  if (workB > 0) {
    ProfScope ps_(c, VSR_PROF_SCORE);
    sp.pass = 1;  // pass B: recompute the scores, write P and partial row sums
    if (c.attn_2cta) launch_tc2<Score2Policy>(c, sp, workB2);
    else launch_tc<ScorePolicy>(c, sp, workB);
  }
  ProfScope ps_(c, VSR_PROF_PV);
  if (c.attn_2cta) launch_tc2<PV2Policy>(c, pp, pv_work2);
  else launch_tc<PVPolicy>(c, pp, pv_work);
}

// ------------------------------------------------------------------------------------------------ resize tables on device
struct DevTaps {
  DevBuf i0, i1, a, w0, w1;
  int src = -1, dst = -1;
  ResizeTaps view() const { return ResizeTaps{i0.as<int>(), i1.as<int>(), a.as<float>(), w0.as<short>(), w1.as<short>()}; }
  void build(int src_n, int dst_n, bool vertical, cudaStream_t s) {
    if (src_n == src && dst_n == dst) return;
    HostTaps t = host_resize_taps(src_n, dst_n, vertical);
    upload(i0, t.i0, s); upload(i1, t.i1, s); upload(a, t.a, s); upload(w0, t.w0, s); upload(w1, t.w1, s);
    src = src_n; dst = dst_n;
  }
};

}  // namespace vsr

// ================================================================================================= engine
using namespace vsr;

struct vsr_sttn {
  Ctx ctx;
  vsr_sttn_config cfg;
  int FH = 0, FW = 0;  // feature map (model/4)
  std::map<std::string, std::vector<float>> host_w;
  std::map<std::string, std::vector<int64_t>> host_shape;
  bool ready = false;
  // layers
  DevBuf stem_w, stem_b;
  ConvLayer enc2, enc3, enc4, dec0, dec2, dec4, dec6;
  ConvLayer qkv[8], outl[8], ff0[8], ff1[8];
  // activations
  DevBuf strips, rgb8, e1, e2s, e3, feats16, feats32, xw16, xw32, qkvb, att16, ffn16, up1, d1, d2, up2, d3, comps,
      mask_d, msmall;  // msmall: sttn-det resized mask [model_h, model_w]
  AttnWorkspace attn;
  // Everything a captured chunk graph reads that depends on the job's geometry (T, strip width, strip rows) lives in a
  // per-geometry record with its own device tables, so that a graph captured for one geometry never sees tables that a
  // later job of another geometry rewrote in place (window schedule / first-visit tables, resize taps).
  struct Geom {
    std::vector<Window> sched;
    std::vector<int> visits_h;
    DevBuf sched_d, visits_d;
    DevTaps pre_x, pre_y, post_x, post_y, mpre_x, mpre_y;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_gen = 0, warm_gen = 0;
    bool warm = false;
    int64_t graph_launches = 0;
    uint64_t last_use = 0;
    ~Geom() {
      if (graph_exec) cudaGraphExecDestroy(graph_exec);
    }
  };
  std::map<std::array<int, 4>, std::unique_ptr<Geom>> geoms;
  Geom* g = nullptr;  // geometry of the job being enqueued
  uint64_t geom_clock = 0;
  // staged job
  int T = 0, H = 0, W = 0, split_h = 0;
  std::vector<std::array<int, 4>> areas;
  std::vector<const uint8_t*> in_ptrs;
  std::vector<uint8_t> mask_h;  // last mask seen (host copy) — the whole-video driver passes the same mask per chunk
  int mask_H = 0;
  int staged_area = -1;
  uint8_t* pinned = nullptr;
  size_t pinned_n = 0;
  // two-deep asynchronous pipeline (vsr_sttn_submit / vsr_sttn_collect): H2D of chunk i+1 and D2H of chunk i-1
  // overlap the kernels of chunk i on a separate copy stream
  struct Slot {
    DevBuf dev_in, dev_out;
    uint8_t* pin_in = nullptr;
    uint8_t* pin_out = nullptr;
    size_t pin_n = 0;
    cudaEvent_t ev_h2d = nullptr, ev_done = nullptr, ev_d2h = nullptr;
    bool busy = false;
    int T = 0, y0 = 0, sh = 0, sw = 0;
    std::vector<const uint8_t*> in_ptrs;
  } slot[2];
  cudaStream_t copy_stream = nullptr;
  int64_t next_ticket = 0;
  // window-level sharding of one chunk over several GPUs (vsr_sttn_shard_*)
  struct ShardJob {
    int rank = 0, world = 1, T = 0;
    int ref_slots = 0, win_slots = 0;         // exchange slots per rank
    std::vector<int> ref_frames;              // reference frames of the chunk: 0, ref_length, 2 ref_length, ...
    std::vector<int> slot_of_window;          // window -> index of its first frame in the prediction exchange buffer
    std::vector<int> own_windows;
    DevBuf refs, preds, visit_tab;            // exchange buffers [world][slots] and the per-frame blend table
    bool active = false;
  } shard;
  // CUDA graph of one chunk's compute (launch-bound inner loop: ~630 kernels + ~200 D2D copies)
  bool use_graph = true;
  size_t window_group = 2;  // windows sharing each launch (VSR_WINDOW_GROUP, 1..2)
  ~vsr_sttn() {
    geoms.clear();
    for (auto& sl : slot) {
      if (sl.pin_in) cudaFreeHost(sl.pin_in);
      if (sl.pin_out) cudaFreeHost(sl.pin_out);
      if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
      if (sl.ev_done) cudaEventDestroy(sl.ev_done);
      if (sl.ev_d2h) cudaEventDestroy(sl.ev_d2h);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (pinned) cudaFreeHost(pinned);
    if (ctx.stream) cudaStreamDestroy(ctx.stream);
  }
};

namespace vsr {

static const float* get_w(vsr_sttn* h, const std::string& name, std::vector<int64_t> shape) {
  auto it = h->host_w.find(name);
  if (it == h->host_w.end()) throw Error(VSR_ERR_STATE, "missing weight tensor: " + name);
  if (h->host_shape[name] != shape) throw Error(VSR_ERR_ARG, "unexpected shape for " + name);
  return it->second.data();
}

static void finalize(vsr_sttn* h) {
  CK(cudaSetDevice(h->ctx.device));
  cudaStream_t s = h->ctx.stream;
  // stem: [64,3,3,3] -> [tap*3+ci][co] fp32
  {
    const float* w = get_w(h, "encoder.0.weight", {64, 3, 3, 3});
    const float* b = get_w(h, "encoder.0.bias", {64});
    std::vector<float> sw(27 * 64);
    for (int co = 0; co < 64; ++co)
      for (int ci = 0; ci < 3; ++ci)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) sw[((ky * 3 + kx) * 3 + ci) * 64 + co] = w[((co * 3 + ci) * 3 + ky) * 3 + kx];
    upload(h->stem_w, sw, s);
    upload(h->stem_b, std::vector<float>(b, b + 64), s);
  }
  pack_conv(h->enc2, get_w(h, "encoder.2.weight", {64, 64, 3, 3}), get_w(h, "encoder.2.bias", {64}), 64, 64, 3, 1, s);
  pack_conv_s2d(h->enc3, get_w(h, "encoder.4.weight", {128, 64, 3, 3}), get_w(h, "encoder.4.bias", {128}), 128, 64, s);
  pack_conv(h->enc4, get_w(h, "encoder.6.weight", {256, 128, 3, 3}), get_w(h, "encoder.6.bias", {256}), 256, 128, 3, 1, s);
  for (int b = 0; b < 8; ++b) {
    const std::string p = "transformer." + std::to_string(b) + ".";
    std::vector<float> w(768 * 256), bias(768);
    const char* names[3] = {"query_embedding", "key_embedding", "value_embedding"};
    for (int j = 0; j < 3; ++j) {
      const float* wj = get_w(h, p + "attention." + names[j] + ".weight", {256, 256, 1, 1});
      const float* bj = get_w(h, p + "attention." + names[j] + ".bias", {256});
      memcpy(w.data() + (size_t)j * 256 * 256, wj, 256 * 256 * sizeof(float));
      memcpy(bias.data() + j * 256, bj, 256 * sizeof(float));
    }
    pack_conv(h->qkv[b], w.data(), bias.data(), 768, 256, 1, 1, s);
    pack_conv(h->outl[b], get_w(h, p + "attention.output_linear.0.weight", {256, 256, 3, 3}),
              get_w(h, p + "attention.output_linear.0.bias", {256}), 256, 256, 3, 1, s);
    pack_conv(h->ff0[b], get_w(h, p + "feed_forward.conv.0.weight", {256, 256, 3, 3}),
              get_w(h, p + "feed_forward.conv.0.bias", {256}), 256, 256, 3, 2, s);
    pack_conv(h->ff1[b], get_w(h, p + "feed_forward.conv.2.weight", {256, 256, 3, 3}),
              get_w(h, p + "feed_forward.conv.2.bias", {256}), 256, 256, 3, 1, s);
  }
  pack_conv(h->dec0, get_w(h, "decoder.0.conv.weight", {128, 256, 3, 3}), get_w(h, "decoder.0.conv.bias", {128}), 128, 256, 3, 1, s);
  pack_conv(h->dec2, get_w(h, "decoder.2.weight", {64, 128, 3, 3}), get_w(h, "decoder.2.bias", {64}), 64, 128, 3, 1, s);
  pack_conv(h->dec4, get_w(h, "decoder.4.conv.weight", {64, 64, 3, 3}), get_w(h, "decoder.4.conv.bias", {64}), 64, 64, 3, 1, s);
  pack_conv(h->dec6, get_w(h, "decoder.6.weight", {3, 64, 3, 3}), get_w(h, "decoder.6.bias", {3}), 3, 64, 3, 1, s);
  h->host_w.clear();
  h->host_shape.clear();
  h->ready = true;
}

// Window schedule / first-visit tables of the geometry record and every work buffer of a T-frame job.
static void prepare_network(vsr_sttn* h, int T) {
  Ctx& c = h->ctx;
  const int MW = h->cfg.model_w, MH = h->cfg.model_h, FH = h->FH, FW = h->FW, C = 256;
  cudaStream_t s = c.stream;
  REQUIRE(T >= 1, "need at least one frame");
  REQUIRE(h->g, "run_network without a geometry record");
  vsr_sttn::Geom& G = *h->g;
  if (G.sched.empty()) {
    G.sched = host_window_schedule(T, h->cfg.neighbor_stride, h->cfg.ref_length);
    // per window: frame_idx[32], first_visit[32]
    std::vector<int> tab(G.sched.size() * 64, 0);
    std::vector<int> visits(T, 0);
    for (size_t wi = 0; wi < G.sched.size(); ++wi) {
      REQUIRE(G.sched[wi].neighbors.size() <= 32, "neighbour window larger than 32 frames");
      for (size_t i = 0; i < G.sched[wi].neighbors.size(); ++i) {
        const int f = G.sched[wi].neighbors[i];
        tab[wi * 64 + i] = f;
        tab[wi * 64 + 32 + i] = visits[f] == 0;
        ++visits[f];
      }
    }
    // one more table: identity frame index + "first visit" everywhere (sharded mode decodes every window into its own slot)
    const size_t ident = tab.size();
    tab.resize(ident + 64, 1);
    for (int i = 0; i < 32; ++i) tab[ident + i] = i;
    upload(G.sched_d, tab, s);
    upload(G.visits_d, visits, s);
    G.visits_h = visits;
  }
  // frames of the largest launch group: ANY window_group windows may share a launch (the sharded path groups a rank's own windows,
  // e.g. windows 4 and 6 of a 50-frame chunk = 30 frames where consecutive pairs never exceed 29), so take the largest ones
  size_t maxw = 0;
  {
    std::vector<size_t> sizes;
    for (auto& w : G.sched) sizes.push_back(w.neighbors.size() + w.refs.size());
    std::sort(sizes.begin(), sizes.end(), std::greater<size_t>());
    for (size_t i = 0; i < sizes.size() && i < h->window_group; ++i) maxw += sizes[i];
  }
  size_t maxn = 0;
  for (auto& w : G.sched) maxn = std::max(maxn, w.neighbors.size());
  const size_t fpix = (size_t)FH * FW;
  h->rgb8.ensure((size_t)T * MH * MW * 4);
  h->e1.ensure((size_t)T * (MH / 2) * (MW / 2) * 64 * 2);
  h->e2s.ensure((size_t)T * fpix * 256 * 2);
  h->e3.ensure((size_t)T * fpix * 128 * 2);
  h->feats16.ensure((size_t)T * fpix * C * 2);
  h->feats32.ensure((size_t)T * fpix * C * 4);
  h->xw16.ensure(maxw * fpix * C * 2);
  h->xw32.ensure(maxw * fpix * C * 4);
  h->qkvb.ensure(maxw * fpix * 3 * C * 2);
  h->att16.ensure(maxw * fpix * C * 2);
  h->ffn16.ensure(maxw * fpix * C * 2);
  h->up1.ensure(maxn * fpix * 4 * C * 2);
  h->d1.ensure(maxn * fpix * 4 * 128 * 2);
  h->d2.ensure(maxn * fpix * 4 * 64 * 2);
  h->up2.ensure(maxn * fpix * 16 * 64 * 2);
  h->d3.ensure(maxn * fpix * 16 * 64 * 2);
  h->comps.ensure((size_t)T * MH * MW * 3 * 4);
  if (h->cfg.mode == 1) h->msmall.ensure((size_t)MH * MW);

}

// A3/A4: crop is already done (strips), cv2-exact down-scale of the T strips (and of the mask strip for sttn-det).
static void run_pre(vsr_sttn* h, int T, int sw, int sh, const uint8_t* mask_strip) {
  Ctx& c = h->ctx;
  const int MW = h->cfg.model_w, MH = h->cfg.model_h;
  cudaStream_t s = c.stream;
  vsr_sttn::Geom& G = *h->g;
  // A3/A4 pre-processing
  ProfScope region(c, VSR_PROF_PREPOST);
  G.pre_x.build(sw, MW, false, s);
  G.pre_y.build(sh, MH, true, s);
  strip_downscale_kernel<<<dim3((MW + 255) / 256, MH, T), 256, 0, s>>>(h->strips.as<uint8_t>(), (size_t)sh * sw * 3, sw, sh,
                                                                        h->rgb8.as<uint8_t>(), MW, MH, T, G.pre_x.view(),
                                                                        G.pre_y.view());
  CK(cudaGetLastError());
  ++c.launches;
  const bool det = h->cfg.mode == 1;
  if (det) {
    h->msmall.ensure((size_t)MH * MW);
    if (mask_strip) {  // D1: the mask strip goes through the same cv2.resize as the frames (sttn_det_inpaint.py:73)
      G.mpre_x.build(sw, MW, false, s);
      G.mpre_y.build(sh, MH, true, s);
      mask_downscale_kernel<<<dim3((MW + 255) / 256, MH), 256, 0, s>>>(mask_strip, sw, sh, h->msmall.as<uint8_t>(), MW, MH,
                                                                       G.mpre_x.view(), G.mpre_y.view());
      CK(cudaGetLastError());
      ++c.launches;
    }
  }
}

// A5: encoder on frames [f0, f0 + n) of the job -> feats16 / feats32 slots of those frames.
static void run_encoder(vsr_sttn* h, int f0, int n) {
  Ctx& c = h->ctx;
  const int MW = h->cfg.model_w, MH = h->cfg.model_h, FH = h->FH, FW = h->FW, C = 256;
  cudaStream_t s = c.stream;
  const uint8_t* det_mask = h->cfg.mode == 1 ? h->msmall.as<uint8_t>() : nullptr;
  const size_t fpix = (size_t)FH * FW, hpix = (size_t)(MH / 2) * (MW / 2);
  const int T = n;
  const uchar4* rgb = h->rgb8.as<uchar4>() + (size_t)f0 * MH * MW;
  __half* e1 = h->e1.as<__half>() + (size_t)f0 * hpix * 64;
  __half* e2s = h->e2s.as<__half>() + (size_t)f0 * fpix * 256;
  __half* e3 = h->e3.as<__half>() + (size_t)f0 * fpix * 128;
  __half* f16 = h->feats16.as<__half>() + (size_t)f0 * fpix * C;
  float* f32 = h->feats32.as<float>() + (size_t)f0 * fpix * C;
  // A5 encoder
  {
    ProfScope ps_(c, VSR_PROF_ENCODER);
    const int total = T * (MH / 2) * (MW / 2);
    stem_conv_kernel<<<(total + 63) / 64, 256, 0, s>>>(rgb, MH, MW, h->stem_w.as<float>(), h->stem_b.as<float>(),
                                                        e1, total, det_mask);
    CK(cudaGetLastError());
    ++c.launches;
    ConvIO io;
    io.in = e1; io.T = T; io.H = MH / 2; io.W = MW / 2;
    io.flags = CONV_LRELU | CONV_S2D_STORE; io.out16 = e2s;
    run_conv(c, h->enc2, io);
    ConvIO io3;
    io3.in = e2s; io3.T = T; io3.H = FH; io3.W = FW;
    io3.flags = CONV_LRELU; io3.out16 = e3;
    run_conv(c, h->enc3, io3);
    ConvIO io4;
    io4.in = e3; io4.T = T; io4.H = FH; io4.W = FW;
    io4.flags = CONV_LRELU; io4.out16 = f16; io4.out32 = f32;
    run_conv(c, h->enc4, io4);
  }
}

// A6-A11 for the windows `wins` of the schedule (all of them, in order, for a whole-chunk job).  preds == nullptr: decode into the
// running composites `comps` with the 0.5 / 0.5 blend of sttn_auto_inpaint.py:159-162 (windows must then come in schedule order);
// preds != nullptr (sharded mode): window w's quantised frames go, unblended, to the frames slot_of[w] ... of `preds`.
static void run_windows(vsr_sttn* h, const std::vector<int>& wins, float* preds = nullptr, const std::vector<int>* slot_of = nullptr) {
  Ctx& c = h->ctx;
  const int MW = h->cfg.model_w, MH = h->cfg.model_h, FH = h->FH, FW = h->FW, C = 256;
  cudaStream_t s = c.stream;
  vsr_sttn::Geom& G = *h->g;
  const bool det = h->cfg.mode == 1;
  const uint8_t* det_mask = det ? h->msmall.as<uint8_t>() : nullptr;
  const size_t fpix = (size_t)FH * FW;
  std::unique_ptr<ProfScope> region;
  // A6-A11 window loop.  The transformer passes of different windows are independent (only the blend
  // into comps is ordered), so `group` consecutive windows share every conv / attention launch: at T=15 a
  // single window is 4.05 waves of 128x256 tiles (81 % wave efficiency), two windows are 7.8 (98 %).
  const size_t f16b = fpix * C * 2, f32b = fpix * C * 4;
  for (size_t w0 = 0; w0 < wins.size(); w0 += h->window_group) {
    const size_t w1 = std::min(wins.size(), w0 + h->window_group);
    std::vector<AttnSegment> segs;
    int Tg = 0;
    region = std::make_unique<ProfScope>(c, VSR_PROF_GATHER);
    for (size_t k = w0; k < w1; ++k) {
      const Window& w = G.sched[wins[k]];
      const int nn = (int)w.neighbors.size();
      // gather feats[neighbor_ids + ref_ids] (sttn_auto_inpaint.py:148)
      CK(cudaMemcpyAsync(h->xw16.as<uint8_t>() + (size_t)Tg * f16b, h->feats16.as<uint8_t>() + (size_t)w.neighbors[0] * f16b,
                         (size_t)nn * f16b, cudaMemcpyDeviceToDevice, s));
      CK(cudaMemcpyAsync(h->xw32.as<uint8_t>() + (size_t)Tg * f32b, h->feats32.as<uint8_t>() + (size_t)w.neighbors[0] * f32b,
                         (size_t)nn * f32b, cudaMemcpyDeviceToDevice, s));
      for (size_t r = 0; r < w.refs.size(); ++r) {
        CK(cudaMemcpyAsync(h->xw16.as<uint8_t>() + (Tg + nn + r) * f16b, h->feats16.as<uint8_t>() + (size_t)w.refs[r] * f16b, f16b,
                           cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(h->xw32.as<uint8_t>() + (Tg + nn + r) * f32b, h->feats32.as<uint8_t>() + (size_t)w.refs[r] * f32b, f32b,
                           cudaMemcpyDeviceToDevice, s));
      }
      segs.push_back({Tg, nn + (int)w.refs.size()});
      Tg += nn + (int)w.refs.size();
    }
    region.reset();
    for (int b = 0; b < 8; ++b) {
      // A7: Q,K,V 1x1 projections in one GEMM (auto_sttn.py:172-174)
      ConvIO q;
      q.in = h->xw16.as<__half>(); q.T = Tg; q.H = FH; q.W = FW; q.out16 = h->qkvb.as<__half>(); q.out16_pitch = 3 * C;
      {
        ProfScope ps_(c, VSR_PROF_QKV);
        run_conv(c, h->qkv[b], q);
      }
      // A8: patch attention, one problem per (window, patch geometry)
      run_attention(c, h->attn, h->qkvb.as<__half>(), 3 * C, 0, C, 2 * C, segs, FH, FW, C, h->cfg.n_patch, h->cfg.patch_w,
                    h->cfg.patch_h, h->att16.as<__half>(), C);
      // output_linear + residual (auto_sttn.py:163-164, 237)
      ConvIO o;
      o.in = h->att16.as<__half>(); o.T = Tg; o.H = FH; o.W = FW; o.flags = CONV_LRELU | CONV_RESIDUAL;
      o.out16 = h->xw16.as<__half>(); o.out32 = h->xw32.as<float>(); o.res32 = h->xw32.as<float>();
      {
        ProfScope ps_(c, VSR_PROF_CONV3_RES);
        run_conv(c, h->outl[b], o);
      }
      // A9: feed forward + residual (auto_sttn.py:215-218, 238)
      ConvIO f0;
      f0.in = h->xw16.as<__half>(); f0.T = Tg; f0.H = FH; f0.W = FW; f0.flags = CONV_LRELU; f0.out16 = h->ffn16.as<__half>();
      {
        ProfScope ps_(c, VSR_PROF_CONV3);
        run_conv(c, h->ff0[b], f0);
      }
      ConvIO f1;
      f1.in = h->ffn16.as<__half>(); f1.T = Tg; f1.H = FH; f1.W = FW; f1.flags = CONV_LRELU | CONV_RESIDUAL;
      f1.out16 = h->xw16.as<__half>(); f1.out32 = h->xw32.as<float>(); f1.res32 = h->xw32.as<float>();
      {
        ProfScope ps_(c, VSR_PROF_CONV3_RES);
        run_conv(c, h->ff1[b], f1);
      }
    }
    // A10 decoder on the neighbour frames only (sttn_auto_inpaint.py:150), window by window in schedule
    // order (the 0.5/0.5 blend of :159-162 is order dependent)
    ProfScope ps_dec(c, VSR_PROF_DECODER);
    for (size_t k = w0; k < w1; ++k) {
      const int wi = wins[k];
      const int nn = (int)G.sched[wi].neighbors.size();
      const __half* xin = h->xw16.as<__half>() + (size_t)segs[k - w0].first * fpix * C;
      size_t total = (size_t)nn * (2 * FH) * (2 * FW) * (C / 8);
      upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(xin, nn, FH, FW, C, h->up1.as<__half>());
      CK(cudaGetLastError());
      ++c.launches;
      ConvIO a;
      a.in = h->up1.as<__half>(); a.T = nn; a.H = 2 * FH; a.W = 2 * FW; a.flags = CONV_LRELU; a.out16 = h->d1.as<__half>();
      run_conv(c, h->dec0, a);
      ConvIO b2;
      b2.in = h->d1.as<__half>(); b2.T = nn; b2.H = 2 * FH; b2.W = 2 * FW; b2.flags = CONV_LRELU; b2.out16 = h->d2.as<__half>();
      run_conv(c, h->dec2, b2);
      total = (size_t)nn * (4 * FH) * (4 * FW) * (64 / 8);
      upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h->d2.as<__half>(), nn, 2 * FH, 2 * FW, 64, h->up2.as<__half>());
      CK(cudaGetLastError());
      ++c.launches;
      ConvIO d;
      d.in = h->up2.as<__half>(); d.T = nn; d.H = MH; d.W = MW; d.flags = CONV_LRELU; d.out16 = h->d3.as<__half>();
      run_conv(c, h->dec4, d);
      // A11: tanh + quantise + blend fused into the last conv's epilogue
      ConvIO e;
      e.in = h->d3.as<__half>(); e.T = nn; e.H = MH; e.W = MW; e.flags = CONV_FINAL;
      if (preds) {
        e.comps = preds + (size_t)(*slot_of)[wi] * MH * MW * 3;
        e.frame_idx = G.sched_d.as<int>() + G.sched.size() * 64;
        e.first_visit = e.frame_idx + 32;
      } else {
        e.comps = h->comps.as<float>();
        e.frame_idx = G.sched_d.as<int>() + wi * 64;
        e.first_visit = G.sched_d.as<int>() + wi * 64 + 32;
      }
      e.det_mask = det_mask;
      e.det_rgb = h->rgb8.as<uchar4>();
      run_conv(c, h->dec6, e);
    }
  }
}

// Encoder + window loop + decoder on T strip frames already on the device as u8 [T, sh, sw, 3] BGR
// (sttn_auto_inpaint.py:122-164).  Leaves comps [T,MH,MW,3] fp32 and visits on the device.
// `mask_strip` (sttn-det only): the strip rows of the full-resolution mask on the device, pitch sw; when null in
// det mode the resized mask is taken from h->msmall as is (vsr_sttn_inpaint_strip_masked).
static void run_network(vsr_sttn* h, int T, int sw, int sh, const uint8_t* mask_strip = nullptr) {
  prepare_network(h, T);
  run_pre(h, T, sw, sh, mask_strip);
  run_encoder(h, 0, T);
  std::vector<int> all(h->g->sched.size());
  for (size_t i = 0; i < all.size(); ++i) all[i] = (int)i;
  run_windows(h, all);
}

// Strided host copies (frame strips <-> pinned staging) on a few threads: one thread moves ~10 GB/s, the
// 104 MB of a 1080p chunk would otherwise cost ~10 ms each way.
template <class F>
static void parallel_for(int n, F&& f) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<unsigned>(8, hw ? hw : 1);
  if (nt > n) nt = n;
  if (nt <= 1) {
    for (int i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&, t] {
      for (int i = t; i < n; i += nt) f(i);
    });
  for (auto& x : th) x.join();
}

static void ensure_pinned(vsr_sttn* h, size_t bytes) {
  if (bytes <= h->pinned_n) return;
  if (h->pinned) CK(cudaFreeHost(h->pinned));
  h->pinned = nullptr;
  h->pinned_n = 0;
  CK(cudaMallocHost(&h->pinned, bytes));
  h->pinned_n = bytes;
}

static void check_ready(vsr_sttn* h) {
  if (!h) throw Error(VSR_ERR_ARG, "null engine");
  if (!h->ready) throw Error(VSR_ERR_STATE, "weights not finalized");
  CK(cudaSetDevice(h->ctx.device));
}

static void sync_stream(vsr_sttn* h) {
  cudaError_t e = cudaStreamSynchronize(h->ctx.stream);
  if (e != cudaSuccess)
    throw Error(VSR_ERR_CUDA, std::string("cudaStreamSynchronize -> ") + cudaGetErrorString(e) + device_error_report());
}

// upload strip k of the staged frames, run, composite, leave strips on device
// `snapshot` (optional): the strip rows of every frame, [T][sh*sw*3], saved before an earlier strip was written back
// into frames that alias the inputs (the reference crops every strip from the untouched frames, sttn_auto_inpaint.py:66-73).
static void stage_area(vsr_sttn* h, int k, const uint8_t* snapshot = nullptr) {
  const int y0 = h->areas[k][0], y1 = h->areas[k][1];
  const int sh = y1 - y0, sw = h->W;
  const size_t sb = (size_t)sh * sw * 3;
  h->strips.ensure(sb * h->T);
  ensure_pinned(h, sb * h->T);
  parallel_for(h->T, [&](int t) {
    memcpy(h->pinned + t * sb, snapshot ? snapshot + t * sb : h->in_ptrs[t] + (size_t)y0 * sw * 3, sb);
  });
  CK(cudaMemcpyAsync(h->strips.p, h->pinned, sb * h->T, cudaMemcpyHostToDevice, h->ctx.stream));
  h->staged_area = k;
}

// The geometry record of a job: created on first use, at most 16 kept (least recently used evicted together with its graph).
static vsr_sttn::Geom& select_geom(vsr_sttn* h, int T, int sw, int y0, int y1) {
  const std::array<int, 4> key{{T, sw, y0, y1}};
  auto it = h->geoms.find(key);
  if (it == h->geoms.end()) {
    if (h->geoms.size() >= 16) {
      auto victim = h->geoms.begin();
      for (auto j = h->geoms.begin(); j != h->geoms.end(); ++j)
        if (j->second->last_use < victim->second->last_use) victim = j;
      CK(cudaStreamSynchronize(h->ctx.stream));  // its tables may still be read by enqueued work
      h->geoms.erase(victim);
    }
    it = h->geoms.emplace(key, std::make_unique<vsr_sttn::Geom>()).first;
  }
  it->second->last_use = ++h->geom_clock;
  h->g = it->second.get();
  return *h->g;
}

static void enqueue_area(vsr_sttn* h, int k) {
  const int y0 = h->areas[k][0], y1 = h->areas[k][1];
  const int sh = y1 - y0, sw = h->W;
  cudaStream_t s = h->ctx.stream;
  const bool det = h->cfg.mode == 1;
  const uint8_t* mask_strip = h->mask_d.as<uint8_t>() + (size_t)y0 * sw;
  vsr_sttn::Geom& G = *h->g;
  run_network(h, h->T, sw, sh, det ? mask_strip : nullptr);
  ProfScope ps_(h->ctx, VSR_PROF_PREPOST);
  G.post_x.build(h->cfg.model_w, sw, false, s);
  G.post_y.build(h->cfg.model_h, sh, true, s);
  // sttn-auto: mask ? comp : frame (sttn_auto_inpaint.py:91);  sttn-det: the whole strip is replaced (sttn_det_inpaint.py:93)
  strip_composite_kernel<<<dim3((sw + 255) / 256, sh, h->T), 256, 0, s>>>(
      h->comps.as<float>(), h->cfg.model_w, h->cfg.model_h, G.visits_d.as<int>(), det ? nullptr : mask_strip, sw,
      h->strips.as<uint8_t>(), (size_t)sh * sw * 3, sw, sh, h->T, G.post_x.view(), G.post_y.view());
  CK(cudaGetLastError());
  ++h->ctx.launches;
}

// First call with a given (T, strip geometry): eager (sizes the workspace, uploads tables).  Second call:
// captured into a CUDA graph.  Later calls: one cudaGraphLaunch.  Any reallocation invalidates the graph; the
// geometry-dependent tables are per record, so graphs of different geometries (A/B sections, ragged batches of
// batch_generator) coexist and stay valid across each other's runs.
static void compute_area(vsr_sttn* h, int k) {
  vsr_sttn::Geom& G = select_geom(h, h->T, h->W, h->areas[k][0], h->areas[k][1]);
  cudaStream_t s = h->ctx.stream;
  if (!h->use_graph) {
    enqueue_area(h, k);
    return;
  }
  if (G.graph_exec && G.graph_gen == g_alloc_generation) {
    CK(cudaGraphLaunch(G.graph_exec, s));
    h->ctx.launches += G.graph_launches;
    return;
  }
  if (G.warm && G.warm_gen == g_alloc_generation) {
    if (G.graph_exec) {
      cudaGraphExecDestroy(G.graph_exec);
      G.graph_exec = nullptr;
    }
    const int64_t l0 = h->ctx.launches;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue_area(h, k);
    } catch (...) {
      cudaStreamEndCapture(s, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CK(cudaStreamEndCapture(s, &g));
    G.graph_launches = h->ctx.launches - l0;
    cudaError_t e = cudaGraphInstantiate(&G.graph_exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
      G.graph_exec = nullptr;
      throw Error(VSR_ERR_CUDA, std::string("cudaGraphInstantiate -> ") + cudaGetErrorString(e));
    }
    G.graph_gen = g_alloc_generation;
    CK(cudaGraphLaunch(G.graph_exec, s));
    return;
  }
  enqueue_area(h, k);
  G.warm = true;
  G.warm_gen = g_alloc_generation;
}

// The direct attention heads write P = exp2(logit - own logit) in fp16: exact unless some key beats the token's own logit by > 2^15.5.
// Their kernel clamps and raises a flag; the result of such a job must not be used: the caller switches the engine to the unfused
// path (vsr_sttn_set_option("attn_direct", 0)) and repeats the call (the Python classes do that).  Call with the stream drained.
static void check_attn_overflow(vsr_sttn* h) {
  if (!h->ctx.attn_direct || !h->attn.overflow.p) return;
  int v = 0;
  CK(cudaMemcpy(&v, h->attn.overflow.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (!v) return;
  CK(cudaMemset(h->attn.overflow.p, 0, sizeof(int)));
  throw Error(VSR_ERR_RANGE, "attention logits beyond the range of the single-pass softmax (attn_direct); repeat with attn_direct = 0");
}

static void fetch_area(vsr_sttn* h, int k, uint8_t* const* out) {
  const int y0 = h->areas[k][0], y1 = h->areas[k][1];
  const int sh = y1 - y0, sw = h->W;
  const size_t sb = (size_t)sh * sw * 3;
  CK(cudaMemcpyAsync(h->pinned, h->strips.p, sb * h->T, cudaMemcpyDeviceToHost, h->ctx.stream));
  sync_stream(h);
  check_attn_overflow(h);
  parallel_for(h->T, [&](int t) { memcpy(out[t] + (size_t)y0 * sw * 3, h->pinned + t * sb, sb); });
}

// Mask analysis shared by the synchronous and the asynchronous entry points.  sttn-auto thresholds the mask
// (cv2.threshold(mask,127,1,BINARY), sttn_auto_inpaint.py:48) and uses split_h = int(W*3/16) (:54); sttn-det
// keeps the raw mask (its resized copy gates the encoder input and the low-res composite) and uses
// int(W*5/18) / int(H*5/9) (sttn_det_inpaint.py:48-51).  Returns true when the mask changed.
static bool analyze_mask(vsr_sttn* h, int H, int W, const uint8_t* mask) {
  const bool det = h->cfg.mode == 1;
  h->split_h = det ? (H > W ? (int)((double)H * 5 / 9) : (int)((double)W * 5 / 18)) : (int)((double)W * 3 / 16);
  const size_t mb = (size_t)H * W;
  if (h->mask_h.size() == mb && h->mask_H == H && memcmp(h->mask_h.data(), mask, mb) == 0) return false;
  h->mask_h.assign(mask, mask + mb);
  h->mask_H = H;
  std::vector<uint8_t> m01(mb);
  if (det) {
    for (size_t i = 0; i < mb; ++i) m01[i] = mask[i] != 0;  // get_inpaint_area_by_mask binarises with `mask > 0`
  } else {
    for (size_t i = 0; i < mb; ++i) m01[i] = mask[i] > 127 ? 1 : 0;
  }
  h->areas = host_inpaint_areas(W, H, h->split_h, m01.data(), 1);
  h->mask_d.ensure(mb);
  CK(cudaMemcpyAsync(h->mask_d.p, det ? mask : m01.data(), mb, cudaMemcpyHostToDevice, h->ctx.stream));
  CK(cudaStreamSynchronize(h->ctx.stream));
  return true;
}

static void stage(vsr_sttn* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask) {
  check_ready(h);
  REQUIRE(T >= 1 && H >= 1 && W >= 1 && frames_in && mask, "bad frame batch");
  h->T = T; h->H = H; h->W = W;
  h->in_ptrs.assign(frames_in, frames_in + T);
  analyze_mask(h, H, W, mask);
  h->staged_area = -1;
  if (!h->areas.empty()) stage_area(h, 0);
}

// ---- asynchronous chunk pipeline ------------------------------------------------------------------
static int64_t submit(vsr_sttn* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask) {
  check_ready(h);
  REQUIRE(T >= 1 && H >= 1 && W >= 1 && frames_in && mask, "bad frame batch");
  const int64_t ticket = h->next_ticket;
  vsr_sttn::Slot& sl = h->slot[ticket & 1];
  if (sl.busy) throw Error(VSR_ERR_STATE, "vsr_sttn_submit: two chunks already in flight, collect one first");
  // mask analysis (cached) — same as the synchronous path
  h->T = T; h->H = H; h->W = W;
  {
    const size_t mb = (size_t)H * W;
    const bool same = h->mask_h.size() == mb && h->mask_H == H && memcmp(h->mask_h.data(), mask, mb) == 0;
    if (!same)  // a new mask re-uploads on the compute stream: only with an empty pipeline
      for (auto& o : h->slot)
        if (o.busy) throw Error(VSR_ERR_STATE, "vsr_sttn_submit: the mask may only change when no chunk is in flight");
    analyze_mask(h, H, W, mask);
  }
  if (h->areas.size() != 1) throw Error(VSR_ERR_STATE, "vsr_sttn_submit handles exactly one strip; use vsr_sttn_inpaint_frames");
  if (!h->copy_stream) CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  if (!sl.ev_h2d) {
    CK(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&sl.ev_d2h, cudaEventDisableTiming));
  }
  const int y0 = h->areas[0][0], sh = h->areas[0][1] - y0, sw = W;
  const size_t sb = (size_t)sh * sw * 3, total = sb * T;
  if (total > sl.pin_n) {
    if (sl.pin_in) CK(cudaFreeHost(sl.pin_in));
    if (sl.pin_out) CK(cudaFreeHost(sl.pin_out));
    sl.pin_in = sl.pin_out = nullptr;
    CK(cudaMallocHost(&sl.pin_in, total));
    CK(cudaMallocHost(&sl.pin_out, total));
    sl.pin_n = total;
  }
  sl.dev_in.ensure(total);
  sl.dev_out.ensure(total);
  h->strips.ensure(total);
  sl.T = T; sl.y0 = y0; sl.sh = sh; sl.sw = sw;
  sl.in_ptrs.assign(frames_in, frames_in + T);
  parallel_for(T, [&](int t) { memcpy(sl.pin_in + t * sb, frames_in[t] + (size_t)y0 * sw * 3, sb); });
  CK(cudaMemcpyAsync(sl.dev_in.p, sl.pin_in, total, cudaMemcpyHostToDevice, h->copy_stream));
  CK(cudaEventRecord(sl.ev_h2d, h->copy_stream));
  CK(cudaStreamWaitEvent(h->ctx.stream, sl.ev_h2d, 0));
  CK(cudaMemcpyAsync(h->strips.p, sl.dev_in.p, total, cudaMemcpyDeviceToDevice, h->ctx.stream));
  h->in_ptrs = sl.in_ptrs;
  h->staged_area = 0;
  compute_area(h, 0);
  CK(cudaMemcpyAsync(sl.dev_out.p, h->strips.p, total, cudaMemcpyDeviceToDevice, h->ctx.stream));
  CK(cudaEventRecord(sl.ev_done, h->ctx.stream));
  CK(cudaStreamWaitEvent(h->copy_stream, sl.ev_done, 0));
  CK(cudaMemcpyAsync(sl.pin_out, sl.dev_out.p, total, cudaMemcpyDeviceToHost, h->copy_stream));
  CK(cudaEventRecord(sl.ev_d2h, h->copy_stream));
  sl.busy = true;
  ++h->next_ticket;
  return ticket;
}

static void collect(vsr_sttn* h, int64_t ticket, uint8_t* const* frames_out) {
  check_ready(h);
  REQUIRE(ticket >= 0 && ticket < h->next_ticket && frames_out, "bad ticket");
  vsr_sttn::Slot& sl = h->slot[ticket & 1];
  if (!sl.busy) throw Error(VSR_ERR_STATE, "vsr_sttn_collect: ticket already collected");
  cudaError_t e = cudaEventSynchronize(sl.ev_d2h);
  if (e != cudaSuccess)
    throw Error(VSR_ERR_CUDA, std::string("cudaEventSynchronize -> ") + cudaGetErrorString(e) + device_error_report());
  try {
    check_attn_overflow(h);
  } catch (...) {
    sl.busy = false;
    throw;
  }
  const size_t sb = (size_t)sl.sh * sl.sw * 3, fb = (size_t)h->H * h->W * 3;
  parallel_for(sl.T, [&](int t) {
    if (frames_out[t] != sl.in_ptrs[t]) memcpy(frames_out[t], sl.in_ptrs[t], fb);
    memcpy(frames_out[t] + (size_t)sl.y0 * sl.sw * 3, sl.pin_out + t * sb, sb);
  });
  sl.busy = false;
}

// ---- window-level sharding of one chunk --------------------------------------------------------------------------
// The windows of the chunk's schedule (sttn_auto_inpaint.py:142-146) are dealt round-robin: window w runs on rank w % world.  A rank
// encodes only the frames its own windows decode (their neighbour ranges) plus the reference frames it is the *home* of (reference
// frame number j, i.e. frame j * ref_length, lives on rank j % world); the encoder features of the reference frames — the only data a
// window needs from outside its own neighbourhood (get_ref_index :107-120) — are exchanged with ONE all-gather, the quantised window
// predictions with a second one, and every rank replays the ordered 0.5 / 0.5 blend (:159-162) on the full set.  The two exchange
// buffers are laid out [world][slots per rank][...]: a rank's region goes out, and the gathered buffer comes back, by device-to-device
// copies (vsr_sttn_copy) between the engine's memory and the caller's NCCL tensors; nothing passes through the host.
static size_t shard_ref_bytes(vsr_sttn* h) { return (size_t)h->FH * h->FW * 256 * (2 + 4); }   // fp16 + fp32 features of one frame
static size_t shard_pred_bytes(vsr_sttn* h) { return (size_t)32 * h->cfg.model_h * h->cfg.model_w * 3 * sizeof(float); }
// a rank's region of the prediction exchange buffer: its window slots (32 frames each) + one more frame whose first int carries the
// rank's softmax-range flag, so that after the all-gather every rank takes the same decision (check_attn_overflow) without another
// collective; the extra slot is a whole frame so that the buffer stays an array of frames
static size_t shard_region_frames(vsr_sttn* h) { return (size_t)h->shard.win_slots * 32 + 1; }
static size_t shard_pred_region(vsr_sttn* h) { return shard_region_frames(h) * (shard_pred_bytes(h) / 32); }

static void shard_begin(vsr_sttn* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask, int rank, int world) {
  REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  stage(h, frames_in, T, H, W, mask);
  REQUIRE(h->areas.size() == 1, "window sharding handles exactly one strip; use vsr_sttn_inpaint_frames");
  vsr_sttn::ShardJob& J = h->shard;
  J.rank = rank; J.world = world; J.T = T;
  const int y0 = h->areas[0][0], y1 = h->areas[0][1];
  vsr_sttn::Geom& G = select_geom(h, T, W, y0, y1);
  prepare_network(h, T);
  const int nw = (int)G.sched.size();
  J.ref_frames.clear();
  for (int f = 0; f < T; f += h->cfg.ref_length) J.ref_frames.push_back(f);
  J.ref_slots = ((int)J.ref_frames.size() + world - 1) / world;
  J.win_slots = (nw + world - 1) / world;
  J.slot_of_window.assign(nw, 0);
  J.own_windows.clear();
  for (int w = 0; w < nw; ++w) {
    J.slot_of_window[w] = (int)((w % world) * shard_region_frames(h) + (size_t)(w / world) * 32);
    if (w % world == rank) J.own_windows.push_back(w);
  }
  J.refs.ensure((size_t)world * J.ref_slots * shard_ref_bytes(h));
  J.preds.ensure((size_t)world * shard_pred_region(h));
  // blend table: the visits of every frame in schedule order
  std::vector<int> tab((size_t)T * 4, 0);
  for (int w = 0; w < nw; ++w)
    for (size_t i = 0; i < G.sched[w].neighbors.size(); ++i) {
      const int f = G.sched[w].neighbors[i];
      REQUIRE(tab[f * 4] < 3, "a frame is decoded by more than three windows");
      tab[f * 4 + 1 + tab[f * 4]++] = J.slot_of_window[w] + (int)i;
    }
  upload(J.visit_tab, tab, h->ctx.stream);
  // frames this rank encodes: neighbours of its windows + its home reference frames, as maximal contiguous runs
  std::vector<char> need(T, 0);
  for (int w : J.own_windows)
    for (int f : G.sched[w].neighbors) need[f] = 1;
  for (size_t j = 0; j < J.ref_frames.size(); ++j)
    if ((int)j % world == rank) need[J.ref_frames[j]] = 1;
  const int sh = y1 - y0;
  const bool det = h->cfg.mode == 1;
  run_pre(h, T, W, sh, det ? h->mask_d.as<uint8_t>() + (size_t)y0 * W : nullptr);
  for (int f = 0; f < T;) {
    if (!need[f]) { ++f; continue; }
    int e = f;
    while (e < T && need[e]) ++e;
    run_encoder(h, f, e - f);
    f = e;
  }
  // pack the home reference frames' features into this rank's region of the exchange buffer
  const size_t fpix = (size_t)h->FH * h->FW, b16 = fpix * 256 * 2, b32 = fpix * 256 * 4;
  for (size_t j = 0; j < J.ref_frames.size(); ++j) {
    if ((int)j % world != rank) continue;
    uint8_t* dst = J.refs.as<uint8_t>() + ((size_t)rank * J.ref_slots + j / world) * (b16 + b32);
    CK(cudaMemcpyAsync(dst, h->feats16.as<uint8_t>() + (size_t)J.ref_frames[j] * b16, b16, cudaMemcpyDeviceToDevice, h->ctx.stream));
    CK(cudaMemcpyAsync(dst + b16, h->feats32.as<uint8_t>() + (size_t)J.ref_frames[j] * b32, b32, cudaMemcpyDeviceToDevice, h->ctx.stream));
  }
  sync_stream(h);   // the exchange runs on the caller's (NCCL) stream
  J.active = true;
}

static void shard_windows(vsr_sttn* h) {
  vsr_sttn::ShardJob& J = h->shard;
  REQUIRE(J.active, "vsr_sttn_shard_windows without vsr_sttn_shard_begin");
  const size_t fpix = (size_t)h->FH * h->FW, b16 = fpix * 256 * 2, b32 = fpix * 256 * 4;
  for (size_t j = 0; j < J.ref_frames.size(); ++j) {   // every reference frame's features, from whoever encoded them
    const uint8_t* src = J.refs.as<uint8_t>() + ((size_t)(j % J.world) * J.ref_slots + j / J.world) * (b16 + b32);
    CK(cudaMemcpyAsync(h->feats16.as<uint8_t>() + (size_t)J.ref_frames[j] * b16, src, b16, cudaMemcpyDeviceToDevice, h->ctx.stream));
    CK(cudaMemcpyAsync(h->feats32.as<uint8_t>() + (size_t)J.ref_frames[j] * b32, src + b16, b32, cudaMemcpyDeviceToDevice, h->ctx.stream));
  }
  run_windows(h, J.own_windows, J.preds.as<float>(), &J.slot_of_window);
  {   // this rank's softmax-range flag travels with its predictions
    uint8_t* tail = J.preds.as<uint8_t>() + (size_t)J.rank * shard_pred_region(h) + (size_t)J.win_slots * shard_pred_bytes(h);
    if (h->attn.overflow.p) CK(cudaMemcpyAsync(tail, h->attn.overflow.p, sizeof(int), cudaMemcpyDeviceToDevice, h->ctx.stream));
    else CK(cudaMemsetAsync(tail, 0, sizeof(int), h->ctx.stream));
  }
  sync_stream(h);
}

static void shard_finish(vsr_sttn* h, uint8_t* const* frames_out) {
  vsr_sttn::ShardJob& J = h->shard;
  REQUIRE(J.active && frames_out, "vsr_sttn_shard_finish without vsr_sttn_shard_begin");
  vsr_sttn::Geom& G = *h->g;
  cudaStream_t s = h->ctx.stream;
  const int MW = h->cfg.model_w, MH = h->cfg.model_h;
  const size_t fe = (size_t)MH * MW * 3;
  blend_preds_kernel<<<dim3((unsigned)((fe + 255) / 256), J.T), 256, 0, s>>>(h->comps.as<float>(), J.preds.as<float>(), J.visit_tab.as<int>(), fe);
  CK(cudaGetLastError());
  ++h->ctx.launches;
  const int y0 = h->areas[0][0], y1 = h->areas[0][1], sh = y1 - y0, sw = h->W;
  const bool det = h->cfg.mode == 1;
  const uint8_t* mask_strip = h->mask_d.as<uint8_t>() + (size_t)y0 * sw;
  G.post_x.build(MW, sw, false, s);
  G.post_y.build(MH, sh, true, s);
  strip_composite_kernel<<<dim3((sw + 255) / 256, sh, J.T), 256, 0, s>>>(h->comps.as<float>(), MW, MH, G.visits_d.as<int>(), det ? nullptr : mask_strip,
                                                                          sw, h->strips.as<uint8_t>(), (size_t)sh * sw * 3, sw, sh, J.T, G.post_x.view(),
                                                                          G.post_y.view());
  CK(cudaGetLastError());
  ++h->ctx.launches;
  // this rank hands back the frames f with f % world == rank (every rank holds all composites; the output is dealt, not replicated)
  const size_t sb = (size_t)sh * sw * 3, fb = (size_t)h->H * h->W * 3;
  ensure_pinned(h, sb * J.T);
  for (int t = J.rank; t < J.T; t += J.world)
    CK(cudaMemcpyAsync(h->pinned + t * sb, h->strips.as<uint8_t>() + t * sb, sb, cudaMemcpyDeviceToHost, s));
  sync_stream(h);
  if (h->ctx.attn_direct) {   // any rank's flag stops every rank (they all see the same gathered flags)
    bool any = false;
    for (int r = 0; r < J.world; ++r) {
      int v = 0;
      CK(cudaMemcpy(&v, J.preds.as<uint8_t>() + (size_t)r * shard_pred_region(h) + (size_t)J.win_slots * shard_pred_bytes(h), sizeof(int),
                    cudaMemcpyDeviceToHost));
      any |= v != 0;
    }
    if (any) {
      if (h->attn.overflow.p) CK(cudaMemset(h->attn.overflow.p, 0, sizeof(int)));
      J.active = false;
      throw Error(VSR_ERR_RANGE, "attention logits beyond the range of the single-pass softmax (attn_direct) on some rank; repeat with attn_direct = 0");
    }
  }
  parallel_for((J.T - J.rank + J.world - 1) / J.world, [&](int k) {
    const int t = J.rank + k * J.world;
    if (frames_out[t] != h->in_ptrs[t]) memcpy(frames_out[t], h->in_ptrs[t], fb);
    memcpy(frames_out[t] + (size_t)y0 * sw * 3, h->pinned + t * sb, sb);
  });
  J.active = false;
}

}  // namespace vsr


// ================================================================================================= graph runtime
// Device-tensor runtime behind the DBNet text detector.  Python (vsr_b200/dbnet.py) compiles the
// reference's Paddle PIR program into a list of calls on these entry points; tensors are NHWC fp16 device buffers.
struct RtLayer {
  enum Kind { DENSE, DENSE_S2, DEPTHWISE, DIRECT, DECONV } kind = DENSE;
  ConvLayer tc;          // DENSE / DENSE_S2
  DevBuf w32, b32;       // DEPTHWISE / DIRECT / DECONV
  std::map<size_t, std::unique_ptr<DevBuf>> s2d;  // DENSE_S2: space-to-depth staging per input size (stable addresses for graphs)
  int cin = 0, cout = 0, cin_pitch = 0, cout_pitch = 0, kh = 0, kw = 0, stride = 1, pad_t = 0, pad_l = 0;
};

struct vsr_rt {
  Ctx ctx;
  std::vector<std::unique_ptr<DevBuf>> bufs;
  std::vector<std::unique_ptr<RtLayer>> layers;
  DevBuf image;    // BGR u8 staging of the pre-processing
  std::vector<std::pair<std::unique_ptr<DevBuf>, std::unique_ptr<DevBuf>>> lama_slots;  // u8 image + mask staging per batch slot (LAMA)
  struct SeLayer {
    DevBuf w1, b1, w2, b2, mean;
    int C = 0, mid = 0, residual = 0;
    float slope = 0.f, offset = 0.f;
  };
  std::vector<std::unique_ptr<SeLayer>> se_layers;
  std::map<std::pair<int, int>, std::unique_ptr<RtLayer>> corr_layers;  // (target pixels, channels) -> 1x1 "conv" whose weights are fmap2
  DevBuf norm_stats;                                                    // instance-norm mean | rstd
  DevBuf frames8;                                                       // u8 frame staging of vsr_rt_pp_frames
  struct FftPlan {
    cufftHandle r2c = 0, c2r = 0;
    std::shared_ptr<DevBuf> re, sp;  // fp32 staging of the real side [H][W][pitch] and of the spectrum [H][W/2+1][C] complex
  };
  std::map<std::array<int, 4>, FftPlan> fft_plans;  // (H, W, C, pitch of the real side)
  void* out_host = nullptr;  // pinned u8 staging of vsr_rt_lama_output
  size_t out_host_bytes = 0;
  DevBuf plane;    // fp32 single-channel staging of vsr_rt_download_channel
  void* plane_host = nullptr;  // pinned mirror of `plane`
  size_t plane_host_bytes = 0;
  DevBuf flag;     // [0] int overflow flag raised by the scaled epilogues, [1] uint absmax bits
  DevTaps px, py;  // resize tables of the pre-processing
  struct Scene {      // scene-cut scores (vsr_rt_scene_*)
    int H = 0, W = 0, dh = 0, dw = 0, mode = 0, have_prev = 0, slot = 0;
    DevTaps tx, ty;
    DevBuf sdiv, hdiv, hsv[2], frame[2], sums;
    uint8_t* pin[2] = {nullptr, nullptr};
    size_t pin_n = 0;
  } scene;
  std::vector<cudaGraphExec_t> graphs;
  bool capturing = false;
  int* overflow() { return flag.as<int>(); }
  ~vsr_rt() {
    for (auto g : graphs)
      if (g) cudaGraphExecDestroy(g);
    if (plane_host) cudaFreeHost(plane_host);
    if (out_host) cudaFreeHost(out_host);
    for (auto& p : scene.pin)
      if (p) cudaFreeHost(p);
    for (auto& kv : fft_plans) {
      if (kv.second.r2c) cufftDestroy(kv.second.r2c);
      if (kv.second.c2r) cufftDestroy(kv.second.c2r);
    }
  }
};

namespace vsr {
static void rt_check(vsr_rt* h) {
  if (!h) throw Error(VSR_ERR_ARG, "null runtime");
  CK(cudaSetDevice(h->ctx.device));
}
static void rt_sync(vsr_rt* h) {
  cudaError_t e = cudaStreamSynchronize(h->ctx.stream);
  if (e != cudaSuccess)
    throw Error(VSR_ERR_CUDA, std::string("cudaStreamSynchronize -> ") + cudaGetErrorString(e) + device_error_report());
}
static unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }
}  // namespace vsr

// ================================================================================================= C ABI
extern "C" {

const char* vsr_last_error(void) { return g_err.c_str(); }
const char* vsr_version(void) { return "vsr_b200 0.1 (sm_100a, tcgen05/TMA)"; }

int vsr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

void vsr_sttn_default_config(vsr_sttn_config* cfg) {
  if (!cfg) return;
  memset(cfg, 0, sizeof(*cfg));
  cfg->model_w = 640;
  cfg->model_h = 120;
  cfg->n_patch = 4;
  const int pw[4] = {80, 32, 10, 5}, ph[4] = {15, 6, 5, 3};
  for (int i = 0; i < 4; ++i) {
    cfg->patch_w[i] = pw[i];
    cfg->patch_h[i] = ph[i];
  }
  cfg->neighbor_stride = 5;
  cfg->ref_length = 10;
}

void vsr_sttn_det_config(vsr_sttn_config* cfg) {
  if (!cfg) return;
  vsr_sttn_default_config(cfg);
  cfg->model_w = 432;
  cfg->model_h = 240;
  const int pw[4] = {108, 36, 18, 9}, ph[4] = {60, 20, 10, 5};
  for (int i = 0; i < 4; ++i) {
    cfg->patch_w[i] = pw[i];
    cfg->patch_h[i] = ph[i];
  }
  cfg->mode = 1;
}

int vsr_sttn_create(vsr_sttn_t** out, int device, const vsr_sttn_config* cfg) {
  return guarded([&] {
    REQUIRE(out, "out pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      throw Error(VSR_ERR_CUDA, "no CUDA device: vsr_b200 has no CPU fallback");
    }
    REQUIRE(device >= 0 && device < n, "device index");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw Error(VSR_ERR_CUDA, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                    "; vsr_b200 kernels are sm_100a only");
    CK(cudaSetDevice(device));
    auto* h = new vsr_sttn();
    h->ctx.device = device;
    h->ctx.sms = prop.multiProcessorCount;
    if (cfg) h->cfg = *cfg; else vsr_sttn_default_config(&h->cfg);
    if (h->cfg.model_w % 8 || h->cfg.model_h % 8 || h->cfg.mode < 0 || h->cfg.mode > 1) {
      delete h;
      throw Error(VSR_ERR_ARG, "model size must be a multiple of 8 and mode 0 or 1");
    }
    h->FW = h->cfg.model_w / 4;
    h->FH = h->cfg.model_h / 4;
    CK(cudaStreamCreateWithFlags(&h->ctx.stream, cudaStreamNonBlocking));
    h->use_graph = !env_flag("VSR_NO_GRAPH", false);
    h->ctx.conv_2cta = env_flag("VSR_CONV_2CTA", true);
    h->ctx.conv_halo = env_flag("VSR_CONV_HALO", true);
    h->ctx.conv_halo_narrow = env_flag("VSR_CONV_HALO_NARROW", false);
    h->ctx.conv_tma_store = env_flag("VSR_CONV_TMA_STORE", true);
    h->ctx.conv_halo_base_off = env_flag("VSR_CONV_HALO_BASEOFF", false) ? 1 : 0;
    h->ctx.conv_cluster = getenv("VSR_CONV_CLUSTER") && atoi(getenv("VSR_CONV_CLUSTER")) == 4 ? 4 : 2;
    h->ctx.attn_2cta = env_flag("VSR_ATTN_2CTA", true);
    h->ctx.attn_fused = env_flag("VSR_ATTN_FUSED", false);
    h->ctx.attn_direct = env_flag("VSR_ATTN_DIRECT", true);
    h->ctx.attn_lpt = env_flag("VSR_ATTN_LPT", true);
    h->ctx.conv_prefetch = env_flag("VSR_CONV_PREFETCH", false);
    if (getenv("VSR_WINDOW_GROUP")) h->window_group = (size_t)std::min(2, std::max(1, atoi(getenv("VSR_WINDOW_GROUP"))));
    *out = h;
  });
}

void vsr_sttn_destroy(vsr_sttn_t* h) {
  if (!h) return;
  cudaSetDevice(h->ctx.device);
  cudaStreamSynchronize(h->ctx.stream);
  delete h;
}

int vsr_sttn_set_weight(vsr_sttn_t* h, const char* name, const float* data, const int64_t* shape, int ndim) {
  return guarded([&] {
    REQUIRE(h && name && data && shape && ndim >= 1 && ndim <= 4, "bad weight tensor");
    size_t n = 1;
    std::vector<int64_t> shp(shape, shape + ndim);
    for (auto d : shp) n *= (size_t)d;
    h->host_w[name].assign(data, data + n);
    h->host_shape[name] = shp;
    h->ready = false;
  });
}

int vsr_sttn_finalize_weights(vsr_sttn_t* h) {
  return guarded([&] {
    REQUIRE(h, "null engine");
    finalize(h);
  });
}

int vsr_sttn_inpaint_strip(vsr_sttn_t* h, const uint8_t* frames_bgr, int T, float* comps_out, int32_t* visits_out) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(frames_bgr && comps_out && T >= 1, "bad arguments");
    REQUIRE(h->cfg.mode == 0, "sttn-det engines take the resized mask: use vsr_sttn_inpaint_strip_masked");
    const int MW = h->cfg.model_w, MH = h->cfg.model_h;
    const size_t sb = (size_t)MH * MW * 3;
    h->strips.ensure(sb * T);
    CK(cudaMemcpyAsync(h->strips.p, frames_bgr, sb * T, cudaMemcpyHostToDevice, h->ctx.stream));
    select_geom(h, T, MW, 0, MH);
    run_network(h, T, MW, MH);
    CK(cudaMemcpyAsync(comps_out, h->comps.p, sb * T * sizeof(float), cudaMemcpyDeviceToHost, h->ctx.stream));
    sync_stream(h);
    check_attn_overflow(h);
    if (visits_out)
      for (int t = 0; t < T; ++t) visits_out[t] = h->g->visits_h[t];
  });
}

int vsr_sttn_inpaint_strip_masked(vsr_sttn_t* h, const uint8_t* frames_bgr, const uint8_t* mask_small, int T, float* comps_out,
                                  int32_t* visits_out) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(h->cfg.mode == 1, "vsr_sttn_inpaint_strip_masked needs an sttn-det engine (mode 1)");
    REQUIRE(frames_bgr && mask_small && comps_out && T >= 1, "bad arguments");
    const int MW = h->cfg.model_w, MH = h->cfg.model_h;
    const size_t sb = (size_t)MH * MW * 3;
    h->strips.ensure(sb * T);
    h->msmall.ensure((size_t)MH * MW);
    CK(cudaMemcpyAsync(h->strips.p, frames_bgr, sb * T, cudaMemcpyHostToDevice, h->ctx.stream));
    CK(cudaMemcpyAsync(h->msmall.p, mask_small, (size_t)MH * MW, cudaMemcpyHostToDevice, h->ctx.stream));
    select_geom(h, T, MW, 0, MH);
    run_network(h, T, MW, MH, nullptr);
    CK(cudaMemcpyAsync(comps_out, h->comps.p, sb * T * sizeof(float), cudaMemcpyDeviceToHost, h->ctx.stream));
    sync_stream(h);
    check_attn_overflow(h);
    if (visits_out)
      for (int t = 0; t < T; ++t) visits_out[t] = h->g->visits_h[t];
  });
}

int vsr_sttn_stage(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask) {
  return guarded([&] { stage(h, frames_in, T, H, W, mask); });
}

int vsr_sttn_compute(vsr_sttn_t* h) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(h->T > 0, "nothing staged");
    if (h->areas.empty()) return;
    if (h->staged_area != 0) stage_area(h, 0);
    compute_area(h, 0);
  });
}

int vsr_sttn_fetch(vsr_sttn_t* h, uint8_t* const* frames_out) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(h->T > 0 && frames_out, "nothing staged");
    const size_t fb = (size_t)h->H * h->W * 3;
    parallel_for(h->T, [&](int t) {
      if (frames_out[t] != h->in_ptrs[t]) memcpy(frames_out[t], h->in_ptrs[t], fb);  // the copy at sttn_auto_inpaint.py:58
    });
    if (h->areas.empty()) return;
    // further strips (rare: several subtitle bands) run back to back.  Strips may overlap, and frames_out may be the input
    // frames themselves: save the input rows of the later strips before the first strip is written back.
    std::vector<std::vector<uint8_t>> snap(h->areas.size());
    for (size_t k = 1; k < h->areas.size(); ++k) {
      const int y0 = h->areas[k][0];
      const size_t sb = (size_t)(h->areas[k][1] - y0) * h->W * 3;
      snap[k].resize(sb * h->T);
      parallel_for(h->T, [&](int t) { memcpy(snap[k].data() + t * sb, h->in_ptrs[t] + (size_t)y0 * h->W * 3, sb); });
    }
    fetch_area(h, 0, frames_out);
    for (size_t k = 1; k < h->areas.size(); ++k) {
      stage_area(h, (int)k, snap[k].data());
      compute_area(h, (int)k);
      fetch_area(h, (int)k, frames_out);
    }
  });
}

int vsr_sttn_inpaint_frames(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask,
                            uint8_t* const* frames_out) {
  int r = vsr_sttn_stage(h, frames_in, T, H, W, mask);
  if (r) return r;
  r = vsr_sttn_compute(h);
  if (r) return r;
  return vsr_sttn_fetch(h, frames_out);
}

int64_t vsr_sttn_submit(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask) {
  int64_t ticket = -1;
  int r = guarded([&] { ticket = submit(h, frames_in, T, H, W, mask); });
  return r ? (int64_t)r : ticket;
}

int vsr_sttn_collect(vsr_sttn_t* h, int64_t ticket, uint8_t* const* frames_out) {
  return guarded([&] { collect(h, ticket, frames_out); });
}

int vsr_sttn_shard_begin(vsr_sttn_t* h, const uint8_t* const* frames_in, int T, int H, int W, const uint8_t* mask, int rank, int world,
                         void** ref_buf, int64_t* ref_region_bytes, void** pred_buf, int64_t* pred_region_bytes) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(ref_buf && ref_region_bytes && pred_buf && pred_region_bytes, "null output");
    shard_begin(h, frames_in, T, H, W, mask, rank, world);
    *ref_buf = h->shard.refs.p;
    *ref_region_bytes = (int64_t)((size_t)h->shard.ref_slots * shard_ref_bytes(h));
    *pred_buf = h->shard.preds.p;
    *pred_region_bytes = (int64_t)shard_pred_region(h);
  });
}
int vsr_sttn_shard_windows(vsr_sttn_t* h) {
  return guarded([&] {
    check_ready(h);
    shard_windows(h);
  });
}
int vsr_sttn_shard_finish(vsr_sttn_t* h, uint8_t* const* frames_out) {
  return guarded([&] {
    check_ready(h);
    shard_finish(h, frames_out);
  });
}

int vsr_sttn_copy(vsr_sttn_t* h, void* dst, const void* src, int64_t bytes) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(dst && src && bytes >= 0, "bad arguments");
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, h->ctx.stream));
    sync_stream(h);
  });
}

int vsr_sttn_set_option(vsr_sttn_t* h, const char* name, int value) {
  return guarded([&] {
    REQUIRE(h && name, "bad arguments");
    const std::string n(name);
    if (n == "attn_direct") h->ctx.attn_direct = value != 0;
    else if (n == "use_graph") h->use_graph = value != 0;
    else throw Error(VSR_ERR_ARG, "unknown option " + n);
    ++g_alloc_generation;   // captured chunk graphs bake the choice in
  });
}

int vsr_sttn_sync(vsr_sttn_t* h) {
  return guarded([&] {
    REQUIRE(h, "null engine");
    CK(cudaSetDevice(h->ctx.device));
    sync_stream(h);
  });
}
void* vsr_sttn_stream(vsr_sttn_t* h) { return h ? (void*)h->ctx.stream : nullptr; }
int64_t vsr_sttn_launch_count(vsr_sttn_t* h) { return h ? h->ctx.launches : 0; }

int vsr_debug_tc_profile(uint64_t* out64, int reset) {
  return guarded([&] {
    REQUIRE(out64, "out pointer");
    unsigned long long hbuf[64];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(hbuf, g_tc_prof, sizeof(hbuf)));
    for (int i = 0; i < 64; ++i) out64[i] = hbuf[i];
    if (reset) {
      memset(hbuf, 0, sizeof(hbuf));
      CK(cudaMemcpyToSymbol(g_tc_prof, hbuf, sizeof(hbuf)));
    }
  });
}

int vsr_sttn_debug_read(vsr_sttn_t* h, const char* name, float* out, int64_t n) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(name && out && n > 0, "bad arguments");
    sync_stream(h);
    const std::string nm(name);
    const DevBuf* b = nullptr;
    bool is_half = true;
    if (nm == "e1") b = &h->e1;
    else if (nm == "e2s") b = &h->e2s;
    else if (nm == "e3") b = &h->e3;
    else if (nm == "feats16") b = &h->feats16;
    else if (nm == "xw16") b = &h->xw16;
    else if (nm == "att16") b = &h->att16;
    else if (nm == "feats32") { b = &h->feats32; is_half = false; }
    else if (nm == "xw32") { b = &h->xw32; is_half = false; }
    else if (nm == "comps") { b = &h->comps; is_half = false; }
    else throw Error(VSR_ERR_ARG, "unknown debug buffer " + nm);
    REQUIRE((size_t)n * (is_half ? 2 : 4) <= b->n, "debug read larger than the buffer");
    if (is_half) {
      std::vector<__half> tmp((size_t)n);
      CK(cudaMemcpy(tmp.data(), b->p, (size_t)n * 2, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < n; ++i) out[i] = __half2float(tmp[i]);
    } else {
      CK(cudaMemcpy(out, b->p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
  });
}

int vsr_sttn_profile(vsr_sttn_t* h, float* ms_by_class, int64_t* scopes_by_class, int n_classes) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(h->T > 0 && !h->areas.empty(), "nothing staged");
    REQUIRE(ms_by_class && scopes_by_class && n_classes >= VSR_PROF_CLASSES, "need VSR_PROF_CLASSES output slots");
    if (h->staged_area != 0) stage_area(h, 0);
    const bool graph = h->use_graph;
    h->use_graph = false;
    h->ctx.prof = true;
    h->ctx.prof_recs.clear();
    try {
      compute_area(h, 0);
      sync_stream(h);
    } catch (...) {
      h->use_graph = graph;
      h->ctx.prof = false;
      throw;
    }
    h->use_graph = graph;
    h->ctx.prof = false;
    for (int i = 0; i < n_classes; ++i) { ms_by_class[i] = 0.f; scopes_by_class[i] = 0; }
    for (auto& r : h->ctx.prof_recs) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, r.a, r.b));
      ms_by_class[r.cls] += ms;
      ++scopes_by_class[r.cls];
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    h->ctx.prof_recs.clear();
  });
}

int vsr_sttn_time_conv(vsr_sttn_t* h, int T, int n, float* ms_out) {
  return guarded([&] {
    check_ready(h);
    REQUIRE(T >= 1 && n >= 1 && ms_out, "bad arguments");
    const size_t fpix = (size_t)h->FH * h->FW;
    h->xw16.ensure((size_t)T * fpix * 256 * 2);
    h->ffn16.ensure((size_t)T * fpix * 256 * 2);
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    ConvIO f0;
    f0.in = h->xw16.as<__half>(); f0.T = T; f0.H = h->FH; f0.W = h->FW; f0.flags = CONV_LRELU; f0.out16 = h->ffn16.as<__half>();
    run_conv(h->ctx, h->ff1[0], f0);  // warm-up
    CK(cudaEventRecord(ev[0], h->ctx.stream));
    for (int i = 0; i < n; ++i) {
      run_conv(h->ctx, h->ff1[i % 8], f0);
      CK(cudaEventRecord(ev[i + 1], h->ctx.stream));
    }
    sync_stream(h);
    for (int i = 0; i < n; ++i) CK(cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]));
    for (auto& e : ev) cudaEventDestroy(e);
  });
}

// ---- integer path ---------------------------------------------------------------------------------
int vsr_create_mask(uint8_t* mask, int H, int W, const int32_t* boxes, int n, int deviation) {
  return guarded([&] {
    REQUIRE(mask && H > 0 && W > 0 && (n == 0 || boxes), "bad arguments");
    host_create_mask(mask, H, W, boxes, n, deviation);
  });
}
int vsr_inpaint_area_by_mask(int W, int H, int hh, const uint8_t* mask, int multiple, int32_t* areas, int max_areas) {
  int count = 0;
  int r = guarded([&] {
    REQUIRE(mask && H > 0 && W > 0 && hh > 0 && areas, "bad arguments");
    auto a = host_inpaint_areas(W, H, hh, mask, multiple);
    REQUIRE((int)a.size() <= max_areas, "areas buffer too small");
    for (size_t i = 0; i < a.size(); ++i)
      for (int j = 0; j < 4; ++j) areas[4 * i + j] = a[i][j];
    count = (int)a.size();
  });
  return r ? r : count;
}
int vsr_batch_sizes(int n_samples, int max_batch_size, int32_t* sizes, int max_sizes) {
  int count = 0;
  int r = guarded([&] {
    auto v = host_batch_sizes(n_samples, max_batch_size);
    REQUIRE((int)v.size() <= max_sizes, "sizes buffer too small");
    for (size_t i = 0; i < v.size(); ++i) sizes[i] = v[i];
    count = (int)v.size();
  });
  return r ? r : count;
}
int vsr_window_schedule(int T, int stride, int ref_length, int32_t* ids, int32_t* n_neighbors, int32_t* n_refs, int max_windows,
                        int max_ids_per_window) {
  int count = 0;
  int r = guarded([&] {
    REQUIRE(T >= 0 && stride >= 1 && ref_length >= 1 && ids && n_neighbors && n_refs, "bad arguments");
    auto s = host_window_schedule(T, stride, ref_length);
    REQUIRE((int)s.size() <= max_windows, "too many windows");
    for (size_t w = 0; w < s.size(); ++w) {
      REQUIRE((int)(s[w].neighbors.size() + s[w].refs.size()) <= max_ids_per_window, "ids buffer too small");
      int j = 0;
      for (int v : s[w].neighbors) ids[w * max_ids_per_window + j++] = v;
      for (int v : s[w].refs) ids[w * max_ids_per_window + j++] = v;
      n_neighbors[w] = (int)s[w].neighbors.size();
      n_refs[w] = (int)s[w].refs.size();
    }
    count = (int)s.size();
  });
  return r ? r : count;
}


// ---- graph runtime ---------------------------------------------------------------------------------------
int vsr_rt_create(vsr_rt_t** out, int device) {
  return guarded([&] {
    REQUIRE(out, "out pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      throw Error(VSR_ERR_CUDA, "no CUDA device: vsr_b200 has no CPU fallback");
    }
    REQUIRE(device >= 0 && device < n, "device index");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) throw Error(VSR_ERR_CUDA, "vsr_b200 kernels are sm_100a only");
    CK(cudaSetDevice(device));
    auto* h = new vsr_rt();
    h->ctx.device = device;
    h->ctx.sms = prop.multiProcessorCount;
    h->ctx.conv_2cta = env_flag("VSR_CONV_2CTA", true);
    h->ctx.direct_conv_smem = env_flag("VSR_DIRECT_CONV_SMEM", true);
    h->ctx.conv_tma_store = false;   // graph runtime: per-thread stores (its convs use scaled epilogues and cropped grids)
    h->ctx.conv_halo = env_flag("VSR_RT_CONV_HALO", false);   // haloed-tile convs in the graph runtime: LAMA / DBNet / RAFT tests pass with it, but it
    h->ctx.conv_halo_narrow = h->ctx.conv_halo;               // measured no faster there (LAMA 184 vs 192 frames/s): off by default
    CK(cudaStreamCreateWithFlags(&h->ctx.stream, cudaStreamNonBlocking));
    h->flag.ensure(16);
    *out = h;
  });
}
void vsr_rt_destroy(vsr_rt_t* h) {
  if (!h) return;
  cudaSetDevice(h->ctx.device);
  cudaStreamSynchronize(h->ctx.stream);
  cudaStreamDestroy(h->ctx.stream);
  delete h;
}
int vsr_rt_alloc(vsr_rt_t* h, int64_t bytes, uint64_t* dev_ptr) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(bytes > 0 && dev_ptr, "bad arguments");
    auto b = std::make_unique<DevBuf>();
    b->ensure((size_t)bytes);
    *dev_ptr = (uint64_t)(uintptr_t)b->p;
    h->bufs.push_back(std::move(b));
  });
}
int vsr_rt_free(vsr_rt_t* h, uint64_t dev_ptr) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(!h->capturing, "vsr_rt_free during graph capture");
    for (size_t i = 0; i < h->bufs.size(); ++i)
      if ((uint64_t)(uintptr_t)h->bufs[i]->p == dev_ptr) {
        rt_sync(h);  // enqueued work may still use it
        h->bufs.erase(h->bufs.begin() + (long)i);
        ++g_alloc_generation;
        return;
      }
    throw Error(VSR_ERR_ARG, "vsr_rt_free: not a pointer returned by vsr_rt_alloc");
  });
}
int vsr_rt_upload(vsr_rt_t* h, uint64_t dev_ptr, const void* host, int64_t bytes) {
  return guarded([&] {
    rt_check(h);
    CK(cudaMemcpyAsync((void*)(uintptr_t)dev_ptr, host, (size_t)bytes, cudaMemcpyHostToDevice, h->ctx.stream));
    rt_sync(h);
  });
}
int vsr_rt_download(vsr_rt_t* h, uint64_t dev_ptr, void* host, int64_t bytes) {
  return guarded([&] {
    rt_check(h);
    CK(cudaMemcpyAsync(host, (const void*)(uintptr_t)dev_ptr, (size_t)bytes, cudaMemcpyDeviceToHost, h->ctx.stream));
    rt_sync(h);
  });
}
int vsr_rt_copy(vsr_rt_t* h, uint64_t dst, uint64_t src, int64_t bytes) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(dst && src && bytes > 0, "bad arguments");
    CK(cudaMemcpyAsync((void*)(uintptr_t)dst, (const void*)(uintptr_t)src, (size_t)bytes, cudaMemcpyDeviceToDevice, h->ctx.stream));
  });
}
int vsr_rt_sync(vsr_rt_t* h) {
  return guarded([&] {
    rt_check(h);
    rt_sync(h);
  });
}
int64_t vsr_rt_launch_count(vsr_rt_t* h) { return h ? h->ctx.launches : 0; }

/* ---- scene-cut scores ---- */
int vsr_rt_scene_begin(vsr_rt_t* h, int H, int W) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(H > 0 && W > 0, "bad frame size");
    vsr_rt::Scene& S = h->scene;
    S.H = H; S.W = W;
    // scene_manager.py:132-149, 929-933: factor = W // 256, size (round(W / f), round(H / f)) with Python's round-half-even
    const int f = W < 256 ? 1 : W / 256;
    auto pyround = [](double v) { return (int)std::nearbyint(v); };   // default rounding mode: to nearest, ties to even
    S.dw = f <= 1 ? W : pyround((double)W / f);
    S.dh = f <= 1 ? H : pyround((double)H / f);
    S.mode = f <= 1 ? 0 : (W == 2 * S.dw && H == 2 * S.dh) ? 2 : 1;
    cudaStream_t s = h->ctx.stream;
    if (S.mode == 1) {
      S.tx.build(W, S.dw, false, s);
      S.ty.build(H, S.dh, true, s);
    }
    if (!S.sdiv.p) {   // OpenCV's RGB2HSV_b tables: saturate_cast<int>((255 << 12) / (1. * i)), saturate_cast<int>((180 << 12) / (6. * i))
      std::vector<int> sd(256, 0), hd(256, 0);
      for (int i = 1; i < 256; ++i) {
        sd[i] = (int)std::nearbyint((255 << 12) / (1.0 * i));
        hd[i] = (int)std::nearbyint((180 << 12) / (6.0 * i));
      }
      upload(S.sdiv, sd, s);
      upload(S.hdiv, hd, s);
    }
    for (auto& b : S.hsv) b.ensure((size_t)S.dh * S.dw * 4);
    for (auto& b : S.frame) b.ensure((size_t)H * W * 3);
    const size_t fb = (size_t)H * W * 3;
    if (fb > S.pin_n) {
      for (auto& p : S.pin) {
        if (p) CK(cudaFreeHost(p));
        p = nullptr;
        CK(cudaMallocHost(&p, fb));
      }
      S.pin_n = fb;
    }
    S.have_prev = 0;
    S.slot = 0;
    rt_sync(h);
  });
}
int vsr_rt_scene_frames(vsr_rt_t* h, const uint8_t* const* frames_bgr, int n, int64_t* sums_out) {
  return guarded([&] {
    rt_check(h);
    vsr_rt::Scene& S = h->scene;
    REQUIRE(S.H > 0 && frames_bgr && sums_out && n >= 0, "vsr_rt_scene_begin first");
    if (n == 0) return;
    cudaStream_t s = h->ctx.stream;
    const size_t fb = (size_t)S.H * S.W * 3;
    S.sums.ensure((size_t)n * 3 * sizeof(unsigned long long));
    CK(cudaMemsetAsync(S.sums.p, 0, (size_t)n * 3 * sizeof(unsigned long long), s));
    cudaEvent_t ev[2];
    for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (int i = 0; i < n; ++i) {
      const int b = i & 1;
      if (i >= 2) CK(cudaEventSynchronize(ev[b]));   // the pinned buffer's previous upload has been consumed
      memcpy(S.pin[b], frames_bgr[i], fb);
      CK(cudaMemcpyAsync(S.frame[b].p, S.pin[b], fb, cudaMemcpyHostToDevice, s));
      CK(cudaEventRecord(ev[b], s));
      const int cur = S.slot, prev = S.slot ^ 1;
      scene_hsv_diff_kernel<<<dim3((S.dw + 255) / 256, S.dh), 256, 0, s>>>(
          S.frame[b].as<uint8_t>(), S.W, S.H, S.dw, S.dh, S.mode, S.tx.view(), S.ty.view(), S.sdiv.as<int>(), S.hdiv.as<int>(),
          S.hsv[prev].as<uchar4>(), S.hsv[cur].as<uchar4>(), S.have_prev, S.sums.as<unsigned long long>() + (size_t)i * 3);
      CK(cudaGetLastError());
      ++h->ctx.launches;
      S.have_prev = 1;
      S.slot ^= 1;
    }
    CK(cudaMemcpyAsync(sums_out, S.sums.p, (size_t)n * 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    rt_sync(h);
    for (auto& e : ev) cudaEventDestroy(e);
  });
}

int vsr_rt_absmax(vsr_rt_t* h, uint64_t dev_ptr, int64_t n_elems, float* out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(out && n_elems > 0 && n_elems % 8 == 0 && !h->capturing, "bad arguments");
    unsigned int* slot = h->flag.as<unsigned int>() + 1;
    cudaStream_t s = h->ctx.stream;
    CK(cudaMemsetAsync(slot, 0, 4, s));
    const size_t n8 = (size_t)n_elems / 8;
    rt_absmax_kernel<<<(unsigned)std::min<size_t>((n8 + 255) / 256, (size_t)h->ctx.sms * 8), 256, 0, s>>>((const __half*)(uintptr_t)dev_ptr, n8, slot);
    CK(cudaGetLastError());
    ++h->ctx.launches;
    unsigned int bits = 0;
    CK(cudaMemcpyAsync(&bits, slot, 4, cudaMemcpyDeviceToHost, s));
    rt_sync(h);
    memcpy(out, &bits, 4);
  });
}

int vsr_rt_pad(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out, int OH, int OW, int top, int left, int reflect) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(cp % 8 == 0 && OH > 0 && OW > 0 && H > 1 && W > 1, "bad arguments");
    const size_t n = (size_t)T * OH * OW * (cp / 8);
    rt_pad_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, T, H, W, cp, (__half*)(uintptr_t)out, OH, OW, top, left,
                                                            reflect);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_zero_upsample2x(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(cp % 8 == 0, "bad arguments");
    const size_t n = (size_t)T * 4 * H * W * (cp / 8);
    rt_zero_upsample2x_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, T, H, W, cp, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_add_slices(vsr_rt_t* h, int relu, uint64_t a, int pitch_a, uint64_t b, int pitch_b, uint64_t out, int pitch_out, int channels,
                      int64_t pixels, float alpha, float beta) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(channels % 8 == 0 && pitch_a % 8 == 0 && pitch_b % 8 == 0 && pitch_out % 8 == 0 && pixels > 0, "8-channel granularity");
    REQUIRE(a % 16 == 0 && b % 16 == 0 && out % 16 == 0, "16-byte aligned slices");
    rt_add_slices_kernel<<<blocks_for((size_t)pixels * (channels / 8)), 256, 0, h->ctx.stream>>>(
        relu, (const __half*)(uintptr_t)a, pitch_a, (const __half*)(uintptr_t)b, pitch_b, (__half*)(uintptr_t)out, pitch_out, channels / 8,
        (size_t)pixels, alpha, beta, h->overflow());
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_hswish_affine(vsr_rt_t* h, uint64_t in, uint64_t out, int64_t n_elems, float inv_scale_in, float a, float c) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(in && out && n_elems > 0 && n_elems % 8 == 0, "bad arguments");
    rt_hswish_affine_kernel<<<blocks_for((size_t)n_elems / 8), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, (__half*)(uintptr_t)out,
                                                                                       (size_t)n_elems / 8, inv_scale_in, a, c, h->overflow());
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_se_create(vsr_rt_t* h, const float* w1, const float* b1, const float* w2, const float* b2, int C, int mid, float slope, float offset,
                     int residual, int* se_id) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(w1 && b1 && w2 && b2 && se_id && C > 0 && mid > 0 && mid <= 1024, "bad arguments");
    auto L = std::make_unique<vsr_rt::SeLayer>();
    L->C = C; L->mid = mid; L->residual = residual; L->slope = slope; L->offset = offset;
    cudaStream_t s = h->ctx.stream;
    upload(L->w1, std::vector<float>(w1, w1 + (size_t)mid * C), s);
    upload(L->b1, std::vector<float>(b1, b1 + mid), s);
    upload(L->w2, std::vector<float>(w2, w2 + (size_t)C * mid), s);
    upload(L->b2, std::vector<float>(b2, b2 + C), s);
    L->mean.ensure((size_t)((C + 7) / 8 * 8) * 4);
    *se_id = (int)h->se_layers.size();
    h->se_layers.push_back(std::move(L));
  });
}

int vsr_rt_se_gate(vsr_rt_t* h, int se_id, uint64_t x, int64_t pixels, int cp, float inv_scale, uint64_t gate_dev) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(se_id >= 0 && se_id < (int)h->se_layers.size() && x && gate_dev && pixels > 0 && cp % 8 == 0, "bad arguments");
    auto& L = *h->se_layers[se_id];
    REQUIRE(L.C <= cp, "channel count exceeds the tensor pitch");
    cudaStream_t s = h->ctx.stream;
    rt_channel_mean_kernel<<<(L.C + 7) / 8, 256, 0, s>>>((const __half*)(uintptr_t)x, (size_t)pixels, cp, 1.0f / (float)pixels, L.mean.as<float>());
    CK(cudaGetLastError());
    rt_se_fc_kernel<<<1, 256, (size_t)L.mid * 4, s>>>(L.mean.as<float>(), L.w1.as<float>(), L.b1.as<float>(), L.w2.as<float>(), L.b2.as<float>(), L.C,
                                                       L.mid, L.slope, L.offset, inv_scale, L.residual, (float*)(uintptr_t)gate_dev);
    CK(cudaGetLastError());
    h->ctx.launches += 2;
  });
}

// ---- RAFT operators (csrc/pp_ops.cuh) — written against the CPU stand-in, not yet run on a B200 ---------------
int vsr_rt_pp_frames(vsr_rt_t* h, const uint8_t* const* frames_bgr, int T, int H, int W, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(frames_bgr && T > 0 && H > 0 && W > 0 && out && !h->capturing, "bad arguments");
    const size_t fb = (size_t)H * W * 3;
    h->frames8.ensure(fb * T);
    cudaStream_t s = h->ctx.stream;
    for (int t = 0; t < T; ++t) CK(cudaMemcpyAsync(h->frames8.as<uint8_t>() + fb * t, frames_bgr[t], fb, cudaMemcpyHostToDevice, s));
    const size_t px = (size_t)T * H * W;
    pp_frames_to_half_kernel<<<blocks_for(px), 256, 0, s>>>(h->frames8.as<uint8_t>(), px, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_instnorm(vsr_rt_t* h, uint64_t x, int N, int64_t pixels, int cp, int relu, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(x && out && N > 0 && pixels > 0 && cp % 8 == 0 && !h->capturing, "bad arguments");
    h->norm_stats.ensure((size_t)2 * N * cp * 4);
    float* mean = h->norm_stats.as<float>();
    float* rstd = mean + (size_t)N * cp;
    cudaStream_t s = h->ctx.stream;
    pp_instnorm_stats_kernel<<<dim3(cp / 8, N), 256, 0, s>>>((const __half*)(uintptr_t)x, (size_t)pixels, cp, mean, rstd);
    CK(cudaGetLastError());
    const size_t total8 = (size_t)N * pixels * (cp / 8);
    pp_instnorm_apply_kernel<<<blocks_for(total8), 256, 0, s>>>((const __half*)(uintptr_t)x, (size_t)pixels, cp, mean, rstd, relu,
                                                                (__half*)(uintptr_t)out, total8);
    CK(cudaGetLastError());
    h->ctx.launches += 2;
  });
}

int vsr_rt_context_split(vsr_rt_t* h, uint64_t x, int64_t pixels, uint64_t net, int pitch_net, uint64_t inp, int pitch_inp) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(x && net && inp && pixels > 0 && pitch_net % 8 == 0 && pitch_inp % 8 == 0, "bad arguments");
    pp_context_split_kernel<<<blocks_for((size_t)pixels * 32), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)x, (size_t)pixels,
                                                                                      (__half*)(uintptr_t)net, pitch_net, (__half*)(uintptr_t)inp, pitch_inp);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_corr_volume(vsr_rt_t* h, uint64_t fmap1, uint64_t fmap2, int hh, int ww, int C, uint64_t out, int out_pitch) {
  return guarded([&] {
    rt_check(h);
    const int hw = hh * ww;
    REQUIRE(fmap1 && fmap2 && out && hw > 0 && C % 64 == 0 && out_pitch % 8 == 0 && out_pitch >= (hw + 7) / 8 * 8 && !h->capturing, "bad arguments");
    auto& L = h->corr_layers[{hw, C}];
    cudaStream_t s = h->ctx.stream;
    if (!L) {   // all-pairs correlation = 1x1 conv of fmap1 whose [Cout][K] weight matrix IS fmap2 ([pixel][channel], K-major)
      L = std::make_unique<RtLayer>();
      L->kind = RtLayer::DENSE;
      ConvLayer& t = L->tc;
      // Cout rounded up to the 8-channel store granularity: the extra weight rows stay zero (fresh buffers are zero-filled and only hw rows are
      // ever copied in), the extra columns land in the padding of the volume's pitch (>= round8(hw), checked above) — odd map sizes such as
      // 25 x 135 (a 1080-wide portrait strip) would otherwise be refused by the conv launcher
      t.cin = C; t.pitch = 0; t.cout = (hw + 7) / 8 * 8; t.cout_pad = pad_cout(t.cout); t.ntaps = 1; t.K = C; t.bn = t.cout_pad < 256 ? t.cout_pad : 256;
      t.dy[0] = 0; t.dx[0] = 0;
      t.w.ensure((size_t)t.cout_pad * C * 2);
      t.b.ensure((size_t)t.cout_pad * 4);
    }
    CK(cudaMemcpyAsync(L->tc.w.p, (const void*)(uintptr_t)fmap2, (size_t)hw * C * 2, cudaMemcpyDeviceToDevice, s));
    ConvIO io;
    io.in = (const __half*)(uintptr_t)fmap1; io.T = 1; io.H = hh; io.W = ww; io.flags = CONV_SCALED;
    io.out16 = (__half*)(uintptr_t)out; io.out16_pitch = out_pitch; io.out16_coff = 0;
    io.alpha = 1.0f / sqrtf((float)C); io.bias_scale = 0.f; io.overflow = h->overflow();   // corr.py:60: / sqrt(dim)
    run_conv(h->ctx, L->tc, io);
  });
}

int vsr_rt_corr_pool(vsr_rt_t* h, uint64_t in, int64_t rows, int h2, int w2, int pitch_in, uint64_t out, int pitch_out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(in && out && rows > 0 && h2 >= 2 && w2 >= 2 && pitch_in >= h2 * w2 && pitch_out >= (h2 / 2) * (w2 / 2), "bad arguments");
    const size_t n = (size_t)rows * (h2 / 2) * (w2 / 2);
    pp_corr_pool_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, (size_t)rows, h2, w2, pitch_in, (__half*)(uintptr_t)out,
                                                                   pitch_out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_corr_lookup(vsr_rt_t* h, const uint64_t* level_ptr, const int32_t* level_h, const int32_t* level_w, const int32_t* level_pitch, uint64_t flow32,
                       int hh, int ww, int64_t pixels, uint64_t out, int out_pitch) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(level_ptr && level_h && level_w && level_pitch && flow32 && out && pixels > 0 && out_pitch >= 324, "bad arguments");
    CorrLevels lv;
    for (int l = 0; l < 4; ++l) {
      lv.ptr[l] = (const __half*)(uintptr_t)level_ptr[l];
      lv.h[l] = level_h[l]; lv.w[l] = level_w[l]; lv.pitch[l] = level_pitch[l];
    }
    pp_corr_lookup_kernel<<<blocks_for((size_t)pixels * 324), 256, 0, h->ctx.stream>>>(lv, (const float*)(uintptr_t)flow32, hh, ww, (size_t)pixels,
                                                                                     (__half*)(uintptr_t)out, out_pitch);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_gru_rh(vsr_rt_t* h, uint64_t r, int pitch_r, uint64_t hsrc, int pitch_h, uint64_t out, int pitch_out, int64_t pixels) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(r && hsrc && out && pixels > 0 && pitch_r % 8 == 0 && pitch_h % 8 == 0 && pitch_out % 8 == 0, "bad arguments");
    pp_gru_rh_kernel<<<blocks_for((size_t)pixels * 16), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)r, pitch_r, (const __half*)(uintptr_t)hsrc, pitch_h,
                                                                               (__half*)(uintptr_t)out, pitch_out, (size_t)pixels);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_gru_update(vsr_rt_t* h, uint64_t z, int pitch_z, uint64_t q, int pitch_q, uint64_t hio, int pitch_h, int64_t pixels) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(z && q && hio && pixels > 0 && pitch_z % 8 == 0 && pitch_q % 8 == 0 && pitch_h % 8 == 0, "bad arguments");
    pp_gru_update_kernel<<<blocks_for((size_t)pixels * 16), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)z, pitch_z, (const __half*)(uintptr_t)q, pitch_q,
                                                                                   (__half*)(uintptr_t)hio, pitch_h, (size_t)pixels);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_flow_update(vsr_rt_t* h, uint64_t flow32, uint64_t delta, int pitch_delta, uint64_t flow16, uint64_t dst_a, uint64_t dst_b, int pitch_ab,
                       int coff, int64_t pixels, int add) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(flow32 && flow16 && pixels > 0 && (!add || delta), "bad arguments");
    pp_flow_update_kernel<<<blocks_for((size_t)pixels), 256, 0, h->ctx.stream>>>((float*)(uintptr_t)flow32, (const __half*)(uintptr_t)delta, pitch_delta,
                                                                               (__half*)(uintptr_t)flow16, (__half*)(uintptr_t)dst_a, (__half*)(uintptr_t)dst_b,
                                                                               pitch_ab, coff, (size_t)pixels, add);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_convex_upsample(vsr_rt_t* h, uint64_t flow32, uint64_t mask, int pitch_mask, int N, int hh, int ww, uint64_t out32) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(flow32 && mask && out32 && N > 0 && hh > 0 && ww > 0 && pitch_mask >= 576, "bad arguments");
    pp_convex_upsample_kernel<<<blocks_for((size_t)N * hh * ww * 64), 256, 0, h->ctx.stream>>>((const float*)(uintptr_t)flow32, (const __half*)(uintptr_t)mask,
                                                                                             pitch_mask, N, hh, ww, (float*)(uintptr_t)out32);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_img_prop_step(vsr_rt_t* h, uint64_t prev, uint64_t cur, uint64_t flow_prop, uint64_t flow_check, int H, int W, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(prev && cur && flow_prop && flow_check && out && H > 0 && W > 0, "bad arguments");
    pp_img_prop_step_kernel<<<dim3((W + 255) / 256, H), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)prev, (const __half*)(uintptr_t)cur,
                                                                               (const float*)(uintptr_t)flow_prop, (const float*)(uintptr_t)flow_check, H, W,
                                                                               (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_prop_state(vsr_rt_t* h, uint64_t frames, uint64_t mask_u8, uint64_t prop, int T, int H, int W, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(frames && mask_u8 && out && T > 0 && H > 0 && W > 0, "bad arguments");
    const size_t plane = (size_t)H * W, px = plane * T;
    if (prop)
      pp_state_compose_kernel<<<blocks_for(px), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)frames, (const uint8_t*)(uintptr_t)mask_u8,
                                                                        (const __half*)(uintptr_t)prop, plane, px, (__half*)(uintptr_t)out);
    else
      pp_state_init_kernel<<<blocks_for(px), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)frames, (const uint8_t*)(uintptr_t)mask_u8, plane, px,
                                                                     (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_rfc_input(vsr_rt_t* h, uint64_t flow32, uint64_t mask_u8, int N, int H, int W, int reverse, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(flow32 && mask_u8 && out && N > 0 && H > 0 && W > 0, "bad arguments");
    const size_t plane = (size_t)H * W;
    pp_rfc_input_kernel<<<blocks_for(plane * N), 256, 0, h->ctx.stream>>>((const float*)(uintptr_t)flow32, (const uint8_t*)(uintptr_t)mask_u8, N, plane, reverse,
                                                                        (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_pad_replicate(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out, int OH, int OW, int top, int left) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(in && out && cp % 8 == 0 && OH > 0 && OW > 0, "bad arguments");
    const size_t n = (size_t)T * OH * OW * (cp / 8);
    pp_pad_replicate_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, T, H, W, cp, (__half*)(uintptr_t)out, OH, OW, top, left);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_leaky_relu(vsr_rt_t* h, uint64_t x, int64_t n_elems, float slope) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(x && n_elems > 0 && n_elems % 8 == 0, "bad arguments");
    pp_leaky_relu_kernel<<<blocks_for((size_t)n_elems / 8), 256, 0, h->ctx.stream>>>((__half*)(uintptr_t)x, (size_t)n_elems / 8, slope);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_temporal_taps(vsr_rt_t* h, uint64_t in, int T, int64_t pixels, int cp_in, uint64_t out, int cp_out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(in && out && T > 0 && pixels > 0 && cp_in % 8 == 0 && cp_out >= 3 * cp_in && cp_out % 8 == 0, "bad arguments");
    const size_t n = (size_t)T * pixels * 3 * (cp_in / 8);
    pp_temporal_taps_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, T, (size_t)pixels, cp_in, (__half*)(uintptr_t)out, cp_out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_deform_cols(vsr_rt_t* h, uint64_t xa, int pitch_a, int Ca, uint64_t xb, int pitch_b, int C, int G, uint64_t om, int pitch_om, float max_residue,
                       uint64_t flow32, int H, int W, int64_t pixels, uint64_t cols, int pitch_cols) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(xa && om && cols && C > 0 && G > 0 && C % G == 0 && Ca > 0 && Ca <= C && (Ca == C || xb) && pitch_om >= 27 * G && pitch_cols >= 9 * C && pixels > 0,
            "bad arguments");
    pp_deform_cols_kernel<<<blocks_for((size_t)pixels * G * 9), 256, 0, h->ctx.stream>>>(
        (const __half*)(uintptr_t)xa, pitch_a, Ca, (const __half*)(uintptr_t)xb, pitch_b, C, G, (const __half*)(uintptr_t)om, pitch_om, max_residue,
        (const float*)(uintptr_t)flow32, H, W, (size_t)pixels, (__half*)(uintptr_t)cols, pitch_cols);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_rfc_combine(vsr_rt_t* h, uint64_t pred, int pitch_pred, uint64_t flow32, uint64_t mask_u8, int N, int H, int W, int reverse, uint64_t out32) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(pred && flow32 && mask_u8 && out32 && N > 0 && H > 0 && W > 0, "bad arguments");
    const size_t plane = (size_t)H * W;
    pp_rfc_combine_kernel<<<blocks_for(plane * N), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)pred, pitch_pred, (const float*)(uintptr_t)flow32,
                                                                          (const uint8_t*)(uintptr_t)mask_u8, N, plane, reverse, (float*)(uintptr_t)out32);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_upsample2x_bilinear(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(in && out && T > 0 && H > 1 && W > 1 && cp % 8 == 0, "bad arguments");
    const size_t n = (size_t)T * 4 * H * W * (cp / 8);
    upsample2x_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, T, H, W, cp, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_gen_input(vsr_rt_t* h, uint64_t state, uint64_t mask_u8, uint64_t ids_dev, int n, int H, int W, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(state && mask_u8 && ids_dev && out && n > 0 && H > 0 && W > 0, "bad arguments");
    const size_t plane = (size_t)H * W;
    pp_gen_input_kernel<<<blocks_for(plane * n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)state, (const uint8_t*)(uintptr_t)mask_u8,
                                                                        (const int*)(uintptr_t)ids_dev, n, plane, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_flow_down4(vsr_rt_t* h, uint64_t flow32, uint64_t ids_dev, int n, int H, int W, uint64_t out32) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(flow32 && ids_dev && out32 && n > 0 && H % 4 == 0 && W % 4 == 0, "bad arguments");
    pp_flow_down4_kernel<<<blocks_for((size_t)n * (H / 4) * (W / 4)), 256, 0, h->ctx.stream>>>((const float*)(uintptr_t)flow32, (const int*)(uintptr_t)ids_dev, n,
                                                                                             H, W, (float*)(uintptr_t)out32);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_prop_masks(vsr_rt_t* h, uint64_t gen_in, int n, int H, int W, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(gen_in && out && n > 0 && H % 4 == 0 && W % 4 == 0, "bad arguments");
    pp_prop_masks_kernel<<<blocks_for((size_t)n * (H / 4) * (W / 4)), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)gen_in, n, H, W, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_featprop_cond(vsr_rt_t* h, uint64_t prop, uint64_t cur, int C, uint64_t flow_prop, uint64_t flow_check, uint64_t masks, int H, int W, uint64_t cond,
                         int pitch) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(prop && cur && flow_prop && flow_check && masks && cond && C % 8 == 0 && pitch >= 2 * C + 5, "bad arguments");
    pp_featprop_cond_kernel<<<blocks_for((size_t)H * W * (C / 8)), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)prop, (const __half*)(uintptr_t)cur, C,
                                                                                          (const float*)(uintptr_t)flow_prop, (const float*)(uintptr_t)flow_check,
                                                                                          (const __half*)(uintptr_t)masks, H, W, (__half*)(uintptr_t)cond, pitch);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_write_extra(vsr_rt_t* h, uint64_t src, uint64_t dst, int pitch, int coff, int nch, int64_t pixels) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(src && dst && nch > 0 && nch <= 8 && coff >= 0 && coff + nch <= pitch && pixels > 0, "bad arguments");
    pp_write_extra_kernel<<<blocks_for((size_t)pixels), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)src, (__half*)(uintptr_t)dst, pitch, coff, nch,
                                                                               (size_t)pixels);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_unfold7s3(vsr_rt_t* h, uint64_t in, int n, int hh, int ww, int C, uint64_t out, int pitch, int gelu) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(in && out && n > 0 && C % 8 == 0 && pitch >= 49 * C, "bad arguments");
    const int fh = (hh + 6 - 7) / 3 + 1, fw = (ww + 6 - 7) / 3 + 1;
    pp_unfold7s3_kernel<<<blocks_for((size_t)n * fh * fw * 49 * (C / 8)), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, n, hh, ww, C, fh, fw,
                                                                                                 (__half*)(uintptr_t)out, pitch, gelu);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_fold7s3(vsr_rt_t* h, uint64_t tok, int n, int hh, int ww, int C, int pitch, int normalise, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(tok && out && n > 0 && C % 8 == 0 && pitch >= 49 * C, "bad arguments");
    const int fh = (hh + 6 - 7) / 3 + 1, fw = (ww + 6 - 7) / 3 + 1;
    pp_fold7s3_kernel<<<blocks_for((size_t)n * hh * ww * (C / 8)), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)tok, n, fh, fw, pitch, C, hh, ww, normalise,
                                                                                          (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_layernorm(vsr_rt_t* h, uint64_t x, int64_t tokens, int C, uint64_t gamma_dev, uint64_t beta_dev, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(x && out && gamma_dev && beta_dev && tokens > 0 && C > 0, "bad arguments");
    pp_layernorm_kernel<<<(unsigned)((tokens + 7) / 8), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)x, (size_t)tokens, C, (const float*)(uintptr_t)gamma_dev,
                                                                               (const float*)(uintptr_t)beta_dev, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_pool4(vsr_rt_t* h, uint64_t x, int n, int H, int W, int C, uint64_t w_dev, uint64_t b_dev, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(x && out && w_dev && b_dev && n > 0 && H >= 4 && W >= 4, "bad arguments");
    pp_pool4_kernel<<<blocks_for((size_t)n * (H / 4) * (W / 4) * C), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)x, n, H, W, C, (const float*)(uintptr_t)w_dev,
                                                                                            (const float*)(uintptr_t)b_dev, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_window_attention(vsr_rt_t* h, uint64_t q, uint64_t k, uint64_t v, uint64_t kp, uint64_t vp, int T, int Hn, int Wn, int C, int ph, int pw,
                            uint64_t valid_ind_dev, int n_valid, uint64_t t_ind_dev, int n_tind, uint64_t win_masked_dev, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(q && k && v && kp && vp && out && valid_ind_dev && t_ind_dev && win_masked_dev && T > 0 && Hn % PP_WH == 0 && Wn % PP_WW == 0 && C % 128 == 0,
            "bad arguments");
    dim3 grid((Hn / PP_WH) * (Wn / PP_WW), C / 128, T);
    pp_window_attention_kernel<<<grid, 192, 0, h->ctx.stream>>>((const __half*)(uintptr_t)q, (const __half*)(uintptr_t)k, (const __half*)(uintptr_t)v,
                                                                (const __half*)(uintptr_t)kp, (const __half*)(uintptr_t)vp, T, Hn, Wn, C, ph, pw,
                                                                (const int*)(uintptr_t)valid_ind_dev, n_valid, (const int*)(uintptr_t)t_ind_dev, n_tind,
                                                                (const int*)(uintptr_t)win_masked_dev, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_pred_to_rgb8(vsr_rt_t* h, uint64_t x, int cp, int64_t pixels, uint8_t* host_out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(x && host_out && pixels > 0 && cp >= 3 && !h->capturing, "bad arguments");
    const size_t bytes = (size_t)pixels * 3;
    h->plane.ensure(bytes);
    cudaStream_t s = h->ctx.stream;
    pp_pred_to_rgb8_kernel<<<blocks_for((size_t)pixels), 256, 0, s>>>((const __half*)(uintptr_t)x, cp, (size_t)pixels, h->plane.as<uint8_t>());
    CK(cudaGetLastError());
    ++h->ctx.launches;
    CK(cudaMemcpyAsync(host_out, h->plane.p, bytes, cudaMemcpyDeviceToHost, s));
    rt_sync(h);
  });
}

int vsr_rt_residual_add(vsr_rt_t* h, uint64_t x32, uint64_t y16, uint64_t x16, int64_t n_elems, int init) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(x32 && y16 && x16 && n_elems > 0 && n_elems % 8 == 0, "bad arguments");
    rt_residual_add_kernel<<<blocks_for((size_t)n_elems / 8), 256, 0, h->ctx.stream>>>((float*)(uintptr_t)x32, (const __half*)(uintptr_t)y16,
                                                                                      (__half*)(uintptr_t)x16, (size_t)n_elems / 8, init, h->overflow());
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

namespace vsr {
static void cufft_ck(cufftResult r, const char* what) {
  if (r != CUFFT_SUCCESS) throw Error(VSR_ERR_CUDA, std::string(what) + " -> cuFFT error " + std::to_string((int)r));
}
// plans of the FourierUnit transforms: `batch` contiguous planar signals [H][W] real <-> [H][W/2+1] complex
static vsr_rt::FftPlan& fft_plan(vsr_rt* h, int H, int W, int batch, int planes) {
  const std::array<int, 4> key{H, W, batch, planes};
  auto it = h->fft_plans.find(key);
  if (it != h->fft_plans.end()) return it->second;
  if (h->capturing) throw Error(VSR_ERR_ARG, "FFT plans must be created by a run before graph capture");
  vsr_rt::FftPlan p;
  int n[2] = {H, W};
  const int Wc = W / 2 + 1;
  cufft_ck(cufftPlanMany(&p.r2c, 2, n, nullptr, 1, H * W, nullptr, 1, H * Wc, CUFFT_R2C, batch), "cufftPlanMany(R2C)");
  cufft_ck(cufftPlanMany(&p.c2r, 2, n, nullptr, 1, H * Wc, nullptr, 1, H * W, CUFFT_C2R, batch), "cufftPlanMany(C2R)");
  cufft_ck(cufftSetStream(p.r2c, h->ctx.stream), "cufftSetStream");
  cufft_ck(cufftSetStream(p.c2r, h->ctx.stream), "cufftSetStream");
  p.re = std::make_shared<DevBuf>();
  p.sp = std::make_shared<DevBuf>();
  p.re->ensure((size_t)planes * H * W * 4);
  p.sp->ensure((size_t)batch * H * Wc * 8);
  return h->fft_plans[key] = p;
}
}  // namespace vsr

int vsr_rt_fft_r2c(vsr_rt_t* h, uint64_t in, int T, int H, int W, int C, int cp_in, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(T > 0 && H > 0 && W > 1 && C > 0 && C <= cp_in && cp_in % 8 == 0 && (2 * C) % 8 == 0, "bad arguments");
    REQUIRE(T == 1 || C == cp_in, "batched FourierUnit needs a dense channel pitch");
    const int Wc = W / 2 + 1;
    auto& plan = fft_plan(h, H, W, T * C, T * cp_in);
    cudaStream_t s = h->ctx.stream;
    const size_t P = (size_t)H * W, Pc = (size_t)H * Wc;
    rt_nhwc_to_planar_kernel<1><<<dim3((unsigned)((P + 31) / 32), (cp_in + 31) / 32, T), 256, 0, s>>>((const __half*)(uintptr_t)in, plan.re->as<float>(), P,
                                                                                                     cp_in);
    CK(cudaGetLastError());
    cufft_ck(cufftExecR2C(plan.r2c, plan.re->as<float>(), plan.sp->as<cufftComplex>()), "cufftExecR2C");
    rt_planar_to_nhwc_kernel<2><<<dim3((unsigned)((Pc + 31) / 32), (C + 31) / 32, T), 256, 0, s>>>(plan.sp->as<float>(), (__half*)(uintptr_t)out, Pc, C,
                                                                                                  1.0f / sqrtf((float)H * (float)W), h->overflow());
    CK(cudaGetLastError());
    h->ctx.launches += 3;
  });
}

int vsr_rt_fft_c2r(vsr_rt_t* h, uint64_t in, int T, int H, int W, int C, uint64_t out, int cp_out) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(T > 0 && H > 0 && W > 1 && C > 0 && C <= cp_out && cp_out % 8 == 0 && (2 * C) % 8 == 0, "bad arguments");
    REQUIRE(T == 1 || C == cp_out, "batched FourierUnit needs a dense channel pitch");
    const int Wc = W / 2 + 1;
    auto& plan = fft_plan(h, H, W, T * C, T * cp_out);
    cudaStream_t s = h->ctx.stream;
    const size_t P = (size_t)H * W, Pc = (size_t)H * Wc;
    rt_nhwc_to_planar_kernel<2><<<dim3((unsigned)((Pc + 31) / 32), (C + 31) / 32, T), 256, 0, s>>>((const __half*)(uintptr_t)in, plan.sp->as<float>(), Pc, C);
    CK(cudaGetLastError());
    cufft_ck(cufftExecC2R(plan.c2r, plan.sp->as<cufftComplex>(), plan.re->as<float>()), "cufftExecC2R");
    // planes C..cp_out-1 of the staging are never written by the transform (zero from allocation): the padding channels stay 0
    rt_planar_to_nhwc_kernel<1><<<dim3((unsigned)((P + 31) / 32), (cp_out + 31) / 32, T), 256, 0, s>>>(plan.re->as<float>(), (__half*)(uintptr_t)out, P,
                                                                                                      cp_out, 1.0f / sqrtf((float)H * (float)W),
                                                                                                      h->overflow());
    CK(cudaGetLastError());
    h->ctx.launches += 3;
  });
}

int vsr_rt_lama_input(vsr_rt_t* h, const uint8_t* img, const uint8_t* mask, int ih, int iw, uint64_t out, int H, int W, int cp, int slot) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(img && mask && ih > 0 && iw > 0 && H >= ih && W >= iw && H - ih < ih && W - iw < iw && cp >= 8 && !h->capturing, "bad arguments");
    REQUIRE(slot >= 0 && slot < 64, "staging slot");
    cudaStream_t s = h->ctx.stream;
    while ((int)h->lama_slots.size() <= slot) h->lama_slots.emplace_back(std::make_unique<DevBuf>(), std::make_unique<DevBuf>());
    DevBuf& dimg = *h->lama_slots[slot].first;
    DevBuf& dmask = *h->lama_slots[slot].second;
    dimg.ensure((size_t)ih * iw * 3);
    dmask.ensure((size_t)ih * iw);
    CK(cudaMemcpyAsync(dimg.p, img, (size_t)ih * iw * 3, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dmask.p, mask, (size_t)ih * iw, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync((void*)(uintptr_t)out, 0, (size_t)H * W * cp * 2, s));
    rt_lama_input_kernel<<<dim3((W + 255) / 256, H), 256, 0, s>>>(dimg.as<uint8_t>(), dmask.as<uint8_t>(), ih, iw, (__half*)(uintptr_t)out, H, W, cp);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_lama_output(vsr_rt_t* h, uint64_t pred, int W, int cp, float inv_scale, int ih, int iw, int slot, uint8_t* out) {
  return guarded([&] {
    rt_check(h);
    const size_t bytes = (size_t)ih * iw * 3;
    REQUIRE(out && ih > 0 && iw > 0 && iw <= W && slot >= 0 && slot < (int)h->lama_slots.size() && !h->capturing, "bad arguments");
    DevBuf& dimg = *h->lama_slots[slot].first;
    DevBuf& dmask = *h->lama_slots[slot].second;
    REQUIRE(dimg.n >= bytes && dmask.n >= (size_t)ih * iw, "vsr_rt_lama_input was not called for this slot");
    cudaStream_t s = h->ctx.stream;
    h->plane.ensure(bytes);
    if (h->out_host_bytes < bytes) {
      if (h->out_host) cudaFreeHost(h->out_host);
      h->out_host = nullptr;
      CK(cudaMallocHost(&h->out_host, bytes));
      h->out_host_bytes = bytes;
    }
    rt_lama_output_kernel<<<dim3((iw + 255) / 256, ih), 256, 0, s>>>((const __half*)(uintptr_t)pred, W, cp, inv_scale, dimg.as<uint8_t>(),
                                                                   dmask.as<uint8_t>(), ih, iw, h->plane.as<uint8_t>());
    CK(cudaGetLastError());
    ++h->ctx.launches;
    CK(cudaMemcpyAsync(h->out_host, h->plane.p, bytes, cudaMemcpyDeviceToHost, s));
    rt_sync(h);
    memcpy(out, h->out_host, bytes);
  });
}

int vsr_rt_download_channel(vsr_rt_t* h, uint64_t dev_ptr, int64_t pixels, int cp, int channel, float mul, float* host) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(host && pixels > 0 && channel >= 0 && channel < cp && !h->capturing, "bad arguments");
    cudaStream_t s = h->ctx.stream;
    h->plane.ensure((size_t)pixels * 4);
    if (h->plane_host_bytes < (size_t)pixels * 4) {
      if (h->plane_host) cudaFreeHost(h->plane_host);
      h->plane_host = nullptr;
      CK(cudaMallocHost(&h->plane_host, (size_t)pixels * 4));
      h->plane_host_bytes = (size_t)pixels * 4;
    }
    rt_extract_channel_kernel<<<blocks_for((size_t)pixels), 256, 0, s>>>((const __half*)(uintptr_t)dev_ptr, (size_t)pixels, cp, channel, mul,
                                                                        h->plane.as<float>());
    CK(cudaGetLastError());
    ++h->ctx.launches;
    CK(cudaMemcpyAsync(h->plane_host, h->plane.p, (size_t)pixels * 4, cudaMemcpyDeviceToHost, s));
    rt_sync(h);
    memcpy(host, h->plane_host, (size_t)pixels * 4);
  });
}

int vsr_rt_overflow(vsr_rt_t* h, int* raised) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(raised && !h->capturing, "bad arguments");
    cudaStream_t s = h->ctx.stream;
    CK(cudaMemcpyAsync(raised, h->overflow(), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemsetAsync(h->overflow(), 0, 4, s));
    rt_sync(h);
  });
}

int vsr_rt_capture_begin(vsr_rt_t* h) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(!h->capturing, "capture already in progress");
    CK(cudaStreamBeginCapture(h->ctx.stream, cudaStreamCaptureModeThreadLocal));
    h->capturing = true;
  });
}

int vsr_rt_capture_end(vsr_rt_t* h, int* graph_id) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(h->capturing && graph_id, "no capture in progress");
    h->capturing = false;
    cudaGraph_t g = nullptr;
    CK(cudaStreamEndCapture(h->ctx.stream, &g));
    cudaGraphExec_t ex = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    CK(e);
    *graph_id = (int)h->graphs.size();
    h->graphs.push_back(ex);
  });
}

int vsr_rt_graph_launch(vsr_rt_t* h, int graph_id) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(graph_id >= 0 && graph_id < (int)h->graphs.size() && h->graphs[graph_id], "graph id");
    CK(cudaGraphLaunch(h->graphs[graph_id], h->ctx.stream));
  });
}

int vsr_rt_graph_destroy(vsr_rt_t* h, int graph_id) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(graph_id >= 0 && graph_id < (int)h->graphs.size(), "graph id");
    if (h->graphs[graph_id]) {
      rt_sync(h);
      cudaGraphExecDestroy(h->graphs[graph_id]);
      h->graphs[graph_id] = nullptr;
    }
  });
}

int vsr_rt_conv_create_split(vsr_rt_t* h, const float* w, const float* bias, int cout, int cin, int cin_pitch, int kh, int kw, int pad_t, int pad_l, int dil,
                             int* layer_id) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(w && layer_id && cout >= 8 && cout % 8 == 0 && cin >= 16 && cin_pitch >= cin && cin_pitch % 8 == 0, "split-weight convs are dense tensor-core convs");
    auto L = std::make_unique<RtLayer>();
    L->cin = cin; L->cout = cout; L->cin_pitch = cin_pitch; L->kh = kh; L->kw = kw; L->stride = 1; L->pad_t = pad_t; L->pad_l = pad_l;
    L->kind = RtLayer::DENSE;
    pack_conv_general(L->tc, w, bias, cout, cin, cin_pitch, kh, kw, dil, pad_t, pad_l, h->ctx.stream, true);
    L->cout_pitch = cout;
    *layer_id = (int)h->layers.size();
    h->layers.push_back(std::move(L));
  });
}

int vsr_rt_conv_create(vsr_rt_t* h, const float* w, const float* bias, int cout, int cin, int cin_pitch, int kh, int kw, int stride,
                       int pad_t, int pad_l, int dil, int groups, int transposed, int* layer_id) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(w && layer_id && cout > 0 && cin > 0 && cin_pitch >= cin && cin_pitch % 8 == 0, "bad conv description");
    auto L = std::make_unique<RtLayer>();
    L->cin = cin; L->cout = cout; L->cin_pitch = cin_pitch; L->kh = kh; L->kw = kw; L->stride = stride; L->pad_t = pad_t; L->pad_l = pad_l;
    cudaStream_t s = h->ctx.stream;
    if (transposed) {  // conv2d_transpose 2x2 stride 2: w is [cin][cout][2][2]
      REQUIRE(kh == 2 && kw == 2 && stride == 2 && groups == 1, "only 2x2 stride-2 transposed convs");
      L->kind = RtLayer::DECONV;
      L->cout_pitch = (cout + 7) / 8 * 8;
      std::vector<float> pw((size_t)4 * cin * L->cout_pitch, 0.f), pb(L->cout_pitch, 0.f);
      for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co)
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx)
              pw[((size_t)(dy * 2 + dx) * cin + ci) * L->cout_pitch + co] = w[(((size_t)ci * cout + co) * 2 + dy) * 2 + dx];
      for (int co = 0; co < cout; ++co) pb[co] = bias ? bias[co] : 0.f;
      upload(L->w32, pw, s);
      upload(L->b32, pb, s);
    } else if (groups > 1) {
      REQUIRE(groups == cin && cout == cin && kh == kw && dil == 1, "only depthwise grouped convs");
      L->kind = RtLayer::DEPTHWISE;
      L->cout_pitch = cin_pitch;
      std::vector<float> pw((size_t)kh * kw * cin_pitch, 0.f), pb(cin_pitch, 0.f);
      for (int c = 0; c < cin; ++c)
        for (int t = 0; t < kh * kw; ++t) pw[(size_t)t * cin_pitch + c] = w[(size_t)c * kh * kw + t];
      for (int c = 0; c < cin; ++c) pb[c] = bias ? bias[c] : 0.f;
      upload(L->w32, pw, s);
      upload(L->b32, pb, s);
    } else if (cin < 16 || cout < 8) {
      L->kind = RtLayer::DIRECT;
      REQUIRE(dil == 1, "direct conv without dilation");
      L->cout_pitch = (cout + 7) / 8 * 8;
      std::vector<float> pw((size_t)kh * kw * cin * L->cout_pitch, 0.f), pb(L->cout_pitch, 0.f);
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
          for (int t = 0; t < kh * kw; ++t) pw[((size_t)t * cin + ci) * L->cout_pitch + co] = w[((size_t)co * cin + ci) * kh * kw + t];
      for (int co = 0; co < cout; ++co) pb[co] = bias ? bias[co] : 0.f;
      upload(L->w32, pw, s);
      upload(L->b32, pb, s);
    } else if (stride == 2) {
      REQUIRE(kh == 3 && kw == 3 && pad_t == 1 && pad_l == 1 && dil == 1, "stride-2 dense convs must be 3x3 pad 1");
      REQUIRE(cin_pitch % 16 == 0 && cout % 8 == 0, "channel alignment");
      L->kind = RtLayer::DENSE_S2;
      pack_conv_s2d_pitch(L->tc, w, bias, cout, cin, cin_pitch, s);
      L->cout_pitch = cout;
    } else {
      REQUIRE(stride == 1 && cout % 8 == 0, "dense convs: stride 1, Cout multiple of 8");
      L->kind = RtLayer::DENSE;
      pack_conv_general(L->tc, w, bias, cout, cin, cin_pitch, kh, kw, dil, pad_t, pad_l, s);
      L->cout_pitch = cout;
    }
    *layer_id = (int)h->layers.size();
    h->layers.push_back(std::move(L));
  });
}

int vsr_rt_conv(vsr_rt_t* h, int layer_id, uint64_t in_ptr, int T, int H, int W, uint64_t out_ptr, int out_pitch, int out_coff, int relu,
                float alpha, float bias_scale) {
  return vsr_rt_conv_ex(h, layer_id, in_ptr, T, H, W, out_ptr, out_pitch, out_coff, relu, alpha, bias_scale, 0, 0, 0, 0);
}

int vsr_rt_conv_ex(vsr_rt_t* h, int layer_id, uint64_t in_ptr, int T, int H, int W, uint64_t out_ptr, int out_pitch, int out_coff, int relu,
                   float alpha, float bias_scale, int crop_t, int crop_l, int out_h, int out_w) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(layer_id >= 0 && layer_id < (int)h->layers.size(), "layer id");
    RtLayer& L = *h->layers[layer_id];
    const __half* in = (const __half*)(uintptr_t)in_ptr;
    __half* out = (__half*)(uintptr_t)out_ptr;
    cudaStream_t s = h->ctx.stream;
    Ctx& c = h->ctx;
    const RtScale sc{alpha, bias_scale, h->overflow()};
    const int tc_flags = (relu ? CONV_RELU : 0) | CONV_SCALED;
    const bool cropped = out_h > 0 || crop_t || crop_l;
    REQUIRE(!cropped || L.kind == RtLayer::DENSE || L.kind == RtLayer::DENSE_S2, "cropped output needs a tensor-core conv");
    REQUIRE(!cropped || (out_h > 0 && out_w > 0 && crop_t >= 0 && crop_l >= 0), "bad crop window");
    switch (L.kind) {
      case RtLayer::DENSE: {
        ConvIO io;
        io.in = in; io.T = T; io.H = H; io.W = W; io.flags = tc_flags; io.out16 = out; io.out16_pitch = out_pitch; io.out16_coff = out_coff;
        io.alpha = alpha; io.bias_scale = bias_scale; io.overflow = sc.overflow;
        io.crop_t = crop_t; io.crop_l = crop_l; io.out_H = out_h; io.out_W = out_w;
        run_conv(c, L.tc, io);
        break;
      }
      case RtLayer::DENSE_S2: {
        REQUIRE(H % 2 == 0 && W % 2 == 0, "stride-2 conv needs even input size");
        const size_t n = (size_t)T * H * W * L.cin_pitch;
        auto& stage = L.s2d[n];
        if (!stage) {
          REQUIRE(!h->capturing, "stride-2 conv staging must be allocated by a run before graph capture");
          stage = std::make_unique<DevBuf>();
          stage->ensure(n * 2);
        }
        rt_space_to_depth_kernel<<<blocks_for(n / 8), 256, 0, s>>>(in, T, H, W, L.cin_pitch, stage->as<__half>());
        CK(cudaGetLastError());
        ++c.launches;
        ConvIO io;
        io.in = stage->as<__half>(); io.T = T; io.H = H / 2; io.W = W / 2; io.flags = tc_flags; io.out16 = out;
        io.out16_pitch = out_pitch; io.out16_coff = out_coff;
        io.alpha = alpha; io.bias_scale = bias_scale; io.overflow = sc.overflow;
        io.crop_t = crop_t; io.crop_l = crop_l; io.out_H = out_h; io.out_W = out_w;
        run_conv(c, L.tc, io);
        break;
      }
      case RtLayer::DEPTHWISE: {
        REQUIRE(out_coff == 0 && out_pitch == L.cin_pitch, "depthwise output must be a plain tensor of the same pitch");
        const int OH = (H + 2 * L.pad_t - L.kh) / L.stride + 1, OW = (W + 2 * L.pad_l - L.kw) / L.stride + 1;
        const size_t n = (size_t)T * OH * OW * (L.cin_pitch / 8);
        rt_depthwise_kernel<<<blocks_for(n), 256, 0, s>>>(in, T, H, W, L.cin_pitch, L.w32.as<float>(), L.b32.as<float>(), L.kh, L.stride, L.pad_t,
                                                          relu, out, OH, OW, sc);
        CK(cudaGetLastError());
        ++c.launches;
        break;
      }
      case RtLayer::DIRECT: {
        REQUIRE(out_coff == 0 && out_pitch >= L.cout_pitch && out_pitch % 8 == 0, "direct conv output pitch");
        const int OH = (H + 2 * L.pad_t - L.kh) / L.stride + 1, OW = (W + 2 * L.pad_l - L.kw) / L.stride + 1;
        const size_t n = (size_t)T * OH * OW * (L.cout_pitch / 8);
        const size_t wbytes = (size_t)L.kh * L.kw * L.cin * L.cout_pitch * sizeof(float);
        if (wbytes <= 96 * 1024 && c.direct_conv_smem) {   // weights in shared memory, two pixels per thread
          static size_t configured = 0;
          if (wbytes > 48 * 1024 && wbytes > configured) {
            CK(cudaFuncSetAttribute(rt_direct_conv_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured = 96 * 1024;
          }
          const size_t n2 = (size_t)T * OH * ((OW + 1) / 2) * (L.cout_pitch / 8);
          rt_direct_conv_smem_kernel<<<blocks_for(n2), 256, wbytes, s>>>(in, T, H, W, L.cin_pitch, L.cin, L.w32.as<float>(), L.b32.as<float>(), L.kh,
                                                                        L.kw, L.stride, L.pad_t, L.pad_l, relu, out, OH, OW, L.cout_pitch, out_pitch, sc);
        } else
          rt_direct_conv_kernel<<<blocks_for(n), 256, 0, s>>>(in, T, H, W, L.cin_pitch, L.cin, L.w32.as<float>(), L.b32.as<float>(), L.kh, L.kw,
                                                              L.stride, L.pad_t, L.pad_l, relu, out, OH, OW, L.cout_pitch, out_pitch, sc);
        CK(cudaGetLastError());
        ++c.launches;
        break;
      }
      case RtLayer::DECONV: {
        REQUIRE(out_coff == 0 && out_pitch >= L.cout_pitch && out_pitch % 8 == 0, "deconv output pitch");
        const size_t n = (size_t)T * 2 * H * 2 * W * (L.cout_pitch / 8);
        rt_deconv2x2_kernel<<<blocks_for(n), 256, 0, s>>>(in, T, H, W, L.cin_pitch, L.cin, L.w32.as<float>(), L.b32.as<float>(), relu, out,
                                                          L.cout_pitch, out_pitch, sc);
        CK(cudaGetLastError());
        ++c.launches;
        break;
      }
    }
  });
}

int vsr_rt_elementwise(vsr_rt_t* h, int op, uint64_t a, uint64_t b, uint64_t out, int64_t n_elems, int cp, uint64_t scale_dev,
                       uint64_t shift_dev, float alpha, float beta) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(n_elems > 0 && n_elems % 8 == 0 && cp % 8 == 0, "element count must be a multiple of 8");
    if (op == RT_AFFINE || op == RT_AFFINE_RELU) REQUIRE(scale_dev && shift_dev, "affine needs device scale and shift [cp] fp32");
    rt_elementwise_kernel<<<blocks_for((size_t)n_elems / 8), 256, 0, h->ctx.stream>>>(
        op, (const __half*)(uintptr_t)a, (const __half*)(uintptr_t)b, (__half*)(uintptr_t)out, (size_t)n_elems / 8, cp,
        (const float*)(uintptr_t)scale_dev, (const float*)(uintptr_t)shift_dev, alpha, beta, h->overflow());
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_upsample_nearest(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, int scale, uint64_t out, int out_pitch, int out_coff) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(scale >= 1 && cp % 8 == 0, "bad arguments");
    const size_t n = (size_t)T * H * scale * W * scale * (cp / 8);
    rt_upsample_nearest_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, T, H, W, cp, scale, (__half*)(uintptr_t)out,
                                                                        out_pitch, out_coff);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_maxpool2x2s1(vsr_rt_t* h, uint64_t in, int T, int H, int W, int cp, uint64_t out) {
  return guarded([&] {
    rt_check(h);
    const size_t n = (size_t)T * H * W * (cp / 8);
    rt_maxpool2x2s1_kernel<<<blocks_for(n), 256, 0, h->ctx.stream>>>((const __half*)(uintptr_t)in, T, H, W, cp, (__half*)(uintptr_t)out);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_copy_channels(vsr_rt_t* h, uint64_t src, int src_pitch, uint64_t dst, int dst_pitch, int dst_coff, int channels, int64_t pixels) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(channels % 8 == 0 && dst_coff % 8 == 0 && src_pitch % 8 == 0 && dst_pitch % 8 == 0, "8-channel granularity");
    rt_copy_channels_kernel<<<blocks_for((size_t)pixels * (channels / 8)), 256, 0, h->ctx.stream>>>(
        (const __half*)(uintptr_t)src, src_pitch, (__half*)(uintptr_t)dst, dst_pitch, dst_coff, channels / 8, (size_t)pixels);
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

int vsr_rt_det_preprocess(vsr_rt_t* h, const uint8_t* bgr, int sh, int sw, uint64_t out, int dh, int dw, int cp) {
  return guarded([&] {
    rt_check(h);
    REQUIRE(bgr && sh > 0 && sw > 0 && dh > 0 && dw > 0 && cp >= 8, "bad arguments");
    cudaStream_t s = h->ctx.stream;
    h->image.ensure((size_t)sh * sw * 3);
    CK(cudaMemcpyAsync(h->image.p, bgr, (size_t)sh * sw * 3, cudaMemcpyHostToDevice, s));
    h->px.build(sw, dw, false, s);
    h->py.build(sh, dh, true, s);
    CK(cudaMemsetAsync((void*)(uintptr_t)out, 0, (size_t)dh * dw * cp * 2, s));
    rt_det_preprocess_kernel<<<dim3((dw + 255) / 256, dh), 256, 0, s>>>(h->image.as<uint8_t>(), sw, sh, (__half*)(uintptr_t)out, dw, dh, cp,
                                                                       h->px.i0.as<int>(), h->px.i1.as<int>(), h->px.w0.as<short>(),
                                                                       h->px.w1.as<short>(), h->py.i0.as<int>(), h->py.i1.as<int>(),
                                                                       h->py.w0.as<short>(), h->py.w1.as<short>());
    CK(cudaGetLastError());
    ++h->ctx.launches;
  });
}

// ---- operator-level entry points -------------------------------------------------------------------
namespace {
struct OpCtx {
  Ctx c;
  OpCtx(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      cudaGetLastError();
      throw Error(VSR_ERR_CUDA, "no CUDA device: vsr_b200 has no CPU fallback");
    }
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) throw Error(VSR_ERR_CUDA, "vsr_b200 kernels are sm_100a only");
    c.device = device;
    c.sms = prop.multiProcessorCount;
    c.conv_2cta = env_flag("VSR_CONV_2CTA", true);
    c.conv_halo = env_flag("VSR_CONV_HALO", true);
    c.conv_halo_narrow = env_flag("VSR_CONV_HALO_NARROW", false);
    c.conv_tma_store = env_flag("VSR_CONV_TMA_STORE", true);
    c.conv_halo_base_off = env_flag("VSR_CONV_HALO_BASEOFF", false) ? 1 : 0;
    c.conv_cluster = getenv("VSR_CONV_CLUSTER") && atoi(getenv("VSR_CONV_CLUSTER")) == 4 ? 4 : 2;
    c.attn_2cta = env_flag("VSR_ATTN_2CTA", true);
    c.attn_fused = env_flag("VSR_ATTN_FUSED", false);
    c.attn_direct = env_flag("VSR_ATTN_DIRECT", true);
    c.attn_lpt = env_flag("VSR_ATTN_LPT", true);
    c.conv_prefetch = env_flag("VSR_CONV_PREFETCH", false);
    CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  }
  ~OpCtx() {
    if (c.stream) cudaStreamDestroy(c.stream);
  }
  void sync() {
    cudaError_t e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess)
      throw Error(VSR_ERR_CUDA, std::string("cudaStreamSynchronize -> ") + cudaGetErrorString(e) + device_error_report());
  }
};
static void to_half_dev(DevBuf& d, const float* src, size_t n, cudaStream_t s) {
  std::vector<__half> hbuf(n);
  for (size_t i = 0; i < n; ++i) hbuf[i] = __float2half_rn(src[i]);
  upload(d, hbuf, s);
}
static void from_half_dev(const __half* dptr, float* dst, size_t n, cudaStream_t s) {
  std::vector<__half> hbuf(n);
  CK(cudaMemcpyAsync(hbuf.data(), dptr, n * 2, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (size_t i = 0; i < n; ++i) dst[i] = __half2float(hbuf[i]);
}
}  // namespace

int vsr_op_resize_u8(int device, const uint8_t* src, int sh, int sw, uint8_t* dst, int dh, int dw) {
  return guarded([&] {
    OpCtx o(device);
    REQUIRE(src && dst && sh > 0 && sw > 0 && dh > 0 && dw > 0, "bad arguments");
    DevBuf s, d;
    DevTaps tx, ty;
    s.ensure((size_t)sh * sw * 3);
    d.ensure((size_t)dh * dw * 4);
    CK(cudaMemcpyAsync(s.p, src, (size_t)sh * sw * 3, cudaMemcpyHostToDevice, o.c.stream));
    tx.build(sw, dw, false, o.c.stream);
    ty.build(sh, dh, true, o.c.stream);
    strip_downscale_kernel<<<dim3((dw + 255) / 256, dh, 1), 256, 0, o.c.stream>>>(s.as<uint8_t>(), (size_t)sh * sw * 3, sw, sh,
                                                                                   d.as<uint8_t>(), dw, dh, 1, tx.view(), ty.view());
    CK(cudaGetLastError());
    std::vector<uint8_t> rgba((size_t)dh * dw * 4);
    CK(cudaMemcpyAsync(rgba.data(), d.p, rgba.size(), cudaMemcpyDeviceToHost, o.c.stream));
    o.sync();
    for (size_t i = 0; i < (size_t)dh * dw; ++i) {  // kernel writes RGB(A); give BGR back like cv2
      dst[3 * i + 0] = rgba[4 * i + 2];
      dst[3 * i + 1] = rgba[4 * i + 1];
      dst[3 * i + 2] = rgba[4 * i + 0];
    }
  });
}

int vsr_op_conv2d(int device, const float* in, int T, int H, int W, int Cin, const float* weight, const float* bias, int Cout,
                  int ksize, int dilation, int flags, const float* res, float* out) {
  return guarded([&] {
    OpCtx o(device);
    REQUIRE(in && weight && out && T > 0 && H > 0 && W > 0, "bad arguments");
    ConvLayer L;
    pack_conv(L, weight, bias, Cout, Cin, ksize, dilation, o.c.stream);
    const size_t npix = (size_t)T * H * W;
    DevBuf din, dout16, dout32, dres;
    to_half_dev(din, in, npix * Cin, o.c.stream);
    dout16.ensure(npix * Cout * 2);
    ConvIO io;
    io.in = din.as<__half>(); io.T = T; io.H = H; io.W = W;
    io.flags = (flags & 1 ? CONV_LRELU : 0);
    io.out16 = dout16.as<__half>();
    if (flags & 2) {
      REQUIRE(res, "residual requested without tensor");
      dres.ensure(npix * Cout * 4);
      dout32.ensure(npix * Cout * 4);
      CK(cudaMemcpyAsync(dres.p, res, npix * Cout * 4, cudaMemcpyHostToDevice, o.c.stream));
      io.flags |= CONV_RESIDUAL;
      io.res32 = dres.as<float>();
      io.out32 = dout32.as<float>();
    }
    run_conv(o.c, L, io);
    o.sync();
    if (flags & 2) {
      CK(cudaMemcpy(out, dout32.p, npix * Cout * 4, cudaMemcpyDeviceToHost));
    } else {
      from_half_dev(dout16.as<__half>(), out, npix * Cout, o.c.stream);
    }
  });
}

int vsr_op_conv2d_s2(int device, const float* in, int T, int H, int W, int Cin, const float* weight, const float* bias, int Cout,
                     int flags, float* out) {
  return guarded([&] {
    OpCtx o(device);
    REQUIRE(in && weight && out && T > 0 && H % 2 == 0 && W % 2 == 0, "bad arguments");
    // host space-to-depth of the input: [T,H,W,Cin] -> [T,H/2,W/2,4*Cin]
    const int OH = H / 2, OW = W / 2;
    std::vector<float> s2d((size_t)T * OH * OW * 4 * Cin);
    for (int t = 0; t < T; ++t)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
          memcpy(&s2d[((((size_t)t * OH + y / 2) * OW + x / 2) * 4 + (y & 1) * 2 + (x & 1)) * Cin],
                 &in[(((size_t)t * H + y) * W + x) * Cin], Cin * sizeof(float));
    ConvLayer L;
    pack_conv_s2d(L, weight, bias, Cout, Cin, o.c.stream);
    const size_t npix = (size_t)T * OH * OW;
    DevBuf din, dout16;
    to_half_dev(din, s2d.data(), s2d.size(), o.c.stream);
    dout16.ensure(npix * Cout * 2);
    ConvIO io;
    io.in = din.as<__half>(); io.T = T; io.H = OH; io.W = OW;
    io.flags = (flags & 1 ? CONV_LRELU : 0);
    io.out16 = dout16.as<__half>();
    run_conv(o.c, L, io);
    o.sync();
    from_half_dev(dout16.as<__half>(), out, npix * Cout, o.c.stream);
  });
}

int vsr_op_patch_attention(int device, const float* q, const float* k, const float* v, int T, int H, int W, int C, int n_patch,
                           const int32_t* pw, const int32_t* ph, float* out) {
  return guarded([&] {
    OpCtx o(device);
    REQUIRE(q && k && v && out && pw && ph, "bad arguments");
    const size_t npix = (size_t)T * H * W;
    std::vector<float> qkv(npix * 3 * C);
    for (size_t p = 0; p < npix; ++p) {
      memcpy(&qkv[p * 3 * C], &q[p * C], C * sizeof(float));
      memcpy(&qkv[p * 3 * C + C], &k[p * C], C * sizeof(float));
      memcpy(&qkv[p * 3 * C + 2 * C], &v[p * C], C * sizeof(float));
    }
    DevBuf dq, dout;
    to_half_dev(dq, qkv.data(), qkv.size(), o.c.stream);
    dout.ensure(npix * C * 2);
    AttnWorkspace ws;
    int pwi[4], phi[4];
    for (int i = 0; i < n_patch && i < 4; ++i) { pwi[i] = pw[i]; phi[i] = ph[i]; }
    run_attention(o.c, ws, dq.as<__half>(), 3 * C, 0, C, 2 * C, std::vector<AttnSegment>{{0, T}}, H, W, C, n_patch, pwi, phi,
                  dout.as<__half>(), C);
    o.sync();
    from_half_dev(dout.as<__half>(), out, npix * C, o.c.stream);
  });
}

int vsr_op_upsample2x(int device, const float* in, int T, int H, int W, int C, float* out) {
  return guarded([&] {
    OpCtx o(device);
    REQUIRE(in && out && C % 8 == 0, "bad arguments");
    const size_t npix = (size_t)T * H * W;
    DevBuf din, dout;
    to_half_dev(din, in, npix * C, o.c.stream);
    dout.ensure(npix * 4 * C * 2);
    const size_t total = npix * 4 * (C / 8);
    upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, o.c.stream>>>(din.as<__half>(), T, H, W, C, dout.as<__half>());
    CK(cudaGetLastError());
    o.sync();
    from_half_dev(dout.as<__half>(), out, npix * 4 * C, o.c.stream);
  });
}

}  // extern "C"
