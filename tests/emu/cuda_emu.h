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

// ---- lockstep execution of one block (kernels with warp shuffles / __syncthreads)
// Default: one FIBER per CUDA thread on the calling OS thread — a hand-rolled x86-64 context switch (callee-saved registers + stack pointer),
// round-robin scheduling, a fiber yields while it waits at a barrier.  Deterministic and ~100x cheaper per synchronisation than OS threads,
// which is what makes the window attention affordable at pipeline sizes.  -DEMU_OS_THREADS selects one std::thread per CUDA thread with
// std::barrier instead: the mode for ThreadSanitizer (real concurrency) and AddressSanitizer (no foreign stacks).
#ifndef EMU_OS_THREADS
extern "C" void emu_switch(void** save_sp, void* load_sp);
__asm__(
    ".text\n.globl emu_switch\n.type emu_switch,@function\nemu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size emu_switch, .-emu_switch\n");

struct EmuBarrier {
  int expected = 0, count = 0;
  unsigned gen = 0;
};
struct EmuFiber {
  void* sp = nullptr;
  unsigned tid = 0;
  bool done = false;
  std::unique_ptr<char[]> stack;
};
struct EmuBlock {
  EmuBarrier block_bar;
  std::vector<EmuBarrier> warp_bar;
  std::vector<float> slots[2];
  std::vector<unsigned char> parity;       // per thread: which slot array its next shuffle uses
  std::vector<EmuFiber> fibers;
  EmuFiber* current = nullptr;
  void* sched_sp = nullptr;
  const std::function<void()>* kernel = nullptr;
};
inline EmuBlock* g_emu_block = nullptr;

static inline void emu_yield() { emu_switch(&g_emu_block->current->sp, g_emu_block->sched_sp); }
static inline void emu_barrier_wait(EmuBarrier& b) {
  const unsigned gen = b.gen;
  if (++b.count == b.expected) { b.count = 0; ++b.gen; return; }
  while (b.gen == gen) emu_yield();
}
static inline void emu_barrier_drop(EmuBarrier& b) {       // a thread that returned no longer takes part
  --b.expected;
  if (b.expected > 0 && b.count == b.expected) { b.count = 0; ++b.gen; }
}
static inline void __syncthreads() { emu_barrier_wait(g_emu_block->block_bar); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  EmuBlock& blk = *g_emu_block;
  const unsigned tid = blk.current->tid;
  const unsigned char p = blk.parity[tid];
  blk.parity[tid] = p ^ 1;                 // alternate slot arrays: a lane can only overwrite a slot after every lane has passed the NEXT barrier
  blk.slots[p][tid] = v;
  emu_barrier_wait(blk.warp_bar[tid >> 5]);
  return blk.slots[p][tid ^ (unsigned)lane_mask];
}
extern "C" inline void emu_fiber_main() {
  EmuBlock& blk = *g_emu_block;
  EmuFiber* self = blk.current;
  (*blk.kernel)();
  emu_barrier_drop(blk.warp_bar[self->tid >> 5]);
  emu_barrier_drop(blk.block_bar);
  self->done = true;
  for (;;) emu_yield();
}

static inline void emu_run_block_lockstep(unsigned threads, const std::function<void()>& kernel) {
  constexpr size_t kStack = 256 * 1024;
  EmuBlock blk;
  blk.block_bar.expected = (int)threads;
  blk.warp_bar.resize((threads + 31) / 32);
  for (unsigned w = 0; w < blk.warp_bar.size(); ++w) blk.warp_bar[w].expected = (int)std::min(32u, threads - 32 * w);
  blk.slots[0].assign(threads, 0.f);
  blk.slots[1].assign(threads, 0.f);
  blk.parity.assign(threads, 0);
  blk.kernel = &kernel;
  blk.fibers.resize(threads);
  for (unsigned t = 0; t < threads; ++t) {
    EmuFiber& f = blk.fibers[t];
    f.tid = t;
    f.stack.reset(new char[kStack]);
    uintptr_t top = ((uintptr_t)f.stack.get() + kStack) & ~(uintptr_t)63;
    void** sp = (void**)top;
    *--sp = nullptr;                       // keeps the entry's stack 16-byte aligned after the `ret` below (as after a call)
    *--sp = (void*)&emu_fiber_main;        // `ret` of the first switch jumps here
    for (int i = 0; i < 6; ++i) *--sp = nullptr;   // rbp rbx r12 r13 r14 r15
    f.sp = sp;
  }
  g_emu_block = &blk;
  for (unsigned alive = threads; alive;) {
    alive = 0;
    for (unsigned t = 0; t < threads; ++t) {
      EmuFiber& f = blk.fibers[t];
      if (f.done) continue;
      blk.current = &f;
      threadIdx = dim3(t);
      emu_switch(&blk.sched_sp, f.sp);
      if (!f.done) ++alive;
    }
  }
  g_emu_block = nullptr;
}
#else
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
static inline void emu_run_block_lockstep(unsigned threads, const std::function<void()>& kernel) {
  const dim3 bi = blockIdx, bd = blockDim, gd = gridDim;
  EmuBlock blk;
  blk.block_bar = std::make_unique<std::barrier<>>(threads);
  for (unsigned w = 0; w < (threads + 31) / 32; ++w) blk.warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32u, threads - 32 * w)));
  blk.slots.assign(threads, 0.f);
  g_emu_block = &blk;
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([&, t] {
      blockIdx = bi; blockDim = bd; gridDim = gd; threadIdx = dim3(t);
      kernel();
      blk.warp_bar[t >> 5]->arrive_and_drop();     // a thread that returned no longer takes part in shuffles / barriers
      blk.block_bar->arrive_and_drop();
    });
  for (auto& th : pool) th.join();
  g_emu_block = nullptr;
}
#endif

// grid / block are 1-D..3-D in the grid and 1-D in the block (all kernels of pp_ops.cuh); lockstep = kernel uses shuffles or __syncthreads
static inline void emu_launch(dim3 grid, unsigned threads, bool lockstep, const std::function<void()>& kernel) {
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = dim3(bx, by, bz); blockDim = dim3(threads); gridDim = grid;
        if (lockstep) emu_run_block_lockstep(threads, kernel);
        else
          for (unsigned t = 0; t < threads; ++t) { threadIdx = dim3(t); kernel(); }
      }
}
static inline unsigned emu_blocks(size_t n) { return (unsigned)((n + 255) / 256); }
