// Minimal host emulation of the CUDA constructs csrc/pp_ops.cuh uses, so that its kernels — the real source, not a transcription — can run
// on the CPU against the numpy stand-in of the runtime (tests/test_pp_kernels_emulated.py).  TEST INFRASTRUCTURE ONLY.
// One OS thread per CUDA thread when a kernel needs warp shuffles / __syncthreads (lockstep through std::barrier), a plain loop otherwise.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __grid_constant__

struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) uint4 { unsigned x, y, z, w; };      // 16-byte aligned like on the device: -fsanitize=alignment then flags a misaligned vector access
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

// ---- IEEE binary16 with round-to-nearest-even, like __float2half_rn
struct __half {
  uint16_t bits = 0;
  __half() = default;
};
static inline __half __float2half_rn(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  __half h;
  if (x >= 0x7f800000u) { h.bits = (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0)); return h; }
  if (x >= 0x477ff000u) { h.bits = (uint16_t)(sign | 0x7c00u); return h; }            // rounds to inf
  if (x < 0x33000001u) { h.bits = (uint16_t)sign; return h; }                          // rounds to zero
  int e = (int)(x >> 23) - 127;
  uint32_t m = (x & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t base;
  if (e < -14) { shift = 13 + (-14 - e); base = 0; } else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7fffffu; }
  uint32_t r = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (r & 1))) ++r;
  h.bits = (uint16_t)(sign | (base + r));
  return h;
}
static inline float __half2float(__half h) {
  const uint32_t sign = (uint32_t)(h.bits & 0x8000u) << 16, e = (h.bits >> 10) & 31, m = h.bits & 0x3ffu;
  uint32_t x;
  if (e == 0) {
    if (m == 0) x = sign;
    else {
      const float v = std::ldexp((float)m, -24);
      std::memcpy(&x, &v, 4);
      x |= sign;
    }
  } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
  else x = sign | ((e + 112) << 23) | (m << 13);
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}
struct alignas(4) __half2 { __half x, y; };
static inline float2 __half22float2(__half2 h) { return float2{__half2float(h.x), __half2float(h.y)}; }
static inline __half2 __floats2half2_rn(float a, float b) { __half2 r; r.x = __float2half_rn(a); r.y = __float2half_rn(b); return r; }

static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
#define __expf(a) std::exp((float)(a))
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
using std::max;
using std::min;

// ---- lockstep execution of one block
struct EmuBlock {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<float> slots;
};
inline EmuBlock* g_emu_block = nullptr;
static inline void __syncthreads() { g_emu_block->block_bar->arrive_and_wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  g_emu_block->slots[tid] = v;
  g_emu_block->warp_bar[warp]->arrive_and_wait();
  const float r = g_emu_block->slots[tid ^ (unsigned)lane_mask];
  g_emu_block->warp_bar[warp]->arrive_and_wait();
  return r;
}

// grid / block are 1-D..3-D in the grid and 1-D in the block (all kernels of pp_ops.cuh); lockstep = kernel uses shuffles or __syncthreads
static inline void emu_launch(dim3 grid, unsigned threads, bool lockstep, const std::function<void()>& kernel) {
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        if (!lockstep) {
          blockIdx = dim3(bx, by, bz); blockDim = dim3(threads); gridDim = grid;
          for (unsigned t = 0; t < threads; ++t) { threadIdx = dim3(t); kernel(); }
          continue;
        }
        EmuBlock blk;
        blk.block_bar = std::make_unique<std::barrier<>>(threads);
        for (unsigned w = 0; w < (threads + 31) / 32; ++w) blk.warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32u, threads - 32 * w)));
        blk.slots.assign(threads, 0.f);
        g_emu_block = &blk;
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < threads; ++t)
          pool.emplace_back([&, t] {
            blockIdx = dim3(bx, by, bz); blockDim = dim3(threads); gridDim = grid; threadIdx = dim3(t);
            kernel();
            blk.warp_bar[t >> 5]->arrive_and_drop();     // a thread that returned no longer takes part in shuffles / barriers
            blk.block_bar->arrive_and_drop();
          });
        for (auto& th : pool) th.join();
        g_emu_block = nullptr;
      }
}
static inline unsigned emu_blocks(size_t n) { return (unsigned)((n + 255) / 256); }
