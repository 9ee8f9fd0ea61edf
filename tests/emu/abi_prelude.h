// Host stand-ins for the pieces of csrc/engine.cu that the extracted vsr_rt_* entry points of the ProPainter path lean on (error plumbing, the
// runtime handle, the few CUDA runtime calls).  tests/emu/make_abi_emu.py pastes the REAL entry points below this prelude.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>
#include <memory>
#include "cuda_emu.h"
#include "../../include/vsr_b200.h"              // the extracted definitions must agree with the shipped prototypes
#include "../../video-subtitle-remover_b200/csrc/pp_ops.cuh"

typedef int cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }

namespace vsr {
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline thread_local std::string g_err;
#define CK(expr) do { if ((expr) != cudaSuccess) throw Error(VSR_ERR_CUDA, #expr); } while (0)
#define REQUIRE(cond, msg) do { if (!(cond)) throw Error(VSR_ERR_ARG, std::string(msg) + " [" #cond "]"); } while (0)
template <class F>
static int guarded(F&& f) {
  try { f(); return VSR_OK; }
  catch (const Error& e) { g_err = e.what(); return e.code; }
  catch (const std::exception& e) { g_err = e.what(); return VSR_ERR_STATE; }
}
struct DevBuf {
  void* p = nullptr;
  size_t n = 0;
  void ensure(size_t bytes) {
    if (bytes <= n) return;
    std::free(p);
    p = std::calloc(bytes + 256, 1);
    n = bytes;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
  ~DevBuf() { std::free(p); }
};
struct Ctx { int device = 0; cudaStream_t stream = 0; long launches = 0; };
static unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }
static inline bool emu_lockstep(const char* kernel) {
  const std::string k(kernel);
  return k == "pp_instnorm_stats_kernel" || k == "pp_layernorm_kernel" || k == "pp_window_attention_kernel";
}
}  // namespace vsr
struct vsr_rt {
  vsr::Ctx ctx;
  vsr::DevBuf frames8, norm_stats, plane;
  std::vector<std::unique_ptr<vsr::DevBuf>> bufs;     // "device" allocations of vsr_rt_alloc below
  bool capturing = false;
};
namespace vsr {
static void rt_check(vsr_rt* h) { if (!h) throw Error(VSR_ERR_ARG, "null runtime"); }
static void rt_sync(vsr_rt*) {}
}
using namespace vsr;
#define EMU_LAUNCH(kernel, grid, threads, ...) emu_launch(dim3(grid), (unsigned)(threads), emu_lockstep(#kernel), [&] { kernel(__VA_ARGS__); })
extern "C" {
// host restatements of the runtime's memory entry points (malloc / memcpy), so that the DEVICE flavour of the operator suite
// (tests/pp_op_cases.py DeviceBackend: alloc, upload, launch, download) can be rehearsed on the CPU as well
int vsr_rt_create(vsr_rt_t** out, int) { *out = new vsr_rt(); return VSR_OK; }
void vsr_rt_destroy(vsr_rt_t* h) { delete h; }
int vsr_rt_alloc(vsr_rt_t* h, int64_t bytes, uint64_t* dev_ptr) {
  return guarded([&] {
    REQUIRE(h && bytes > 0 && dev_ptr, "bad arguments");
    h->bufs.push_back(std::make_unique<DevBuf>());
    h->bufs.back()->ensure((size_t)bytes);
    *dev_ptr = (uint64_t)(uintptr_t)h->bufs.back()->p;
  });
}
int vsr_rt_free(vsr_rt_t* h, uint64_t dev_ptr) {
  return guarded([&] {
    for (size_t i = 0; i < h->bufs.size(); ++i)
      if ((uint64_t)(uintptr_t)h->bufs[i]->p == dev_ptr) { h->bufs.erase(h->bufs.begin() + (long)i); return; }
    throw Error(VSR_ERR_ARG, "vsr_rt_free: not a pointer returned by vsr_rt_alloc");
  });
}
int vsr_rt_upload(vsr_rt_t*, uint64_t dev_ptr, const void* host, int64_t bytes) { std::memcpy((void*)(uintptr_t)dev_ptr, host, (size_t)bytes); return VSR_OK; }
int vsr_rt_download(vsr_rt_t*, uint64_t dev_ptr, void* host, int64_t bytes) { std::memcpy(host, (const void*)(uintptr_t)dev_ptr, (size_t)bytes); return VSR_OK; }
int vsr_rt_copy(vsr_rt_t*, uint64_t dst, uint64_t src, int64_t bytes) { std::memmove((void*)(uintptr_t)dst, (const void*)(uintptr_t)src, (size_t)bytes); return VSR_OK; }
int vsr_rt_sync(vsr_rt_t*) { return VSR_OK; }
int64_t vsr_rt_launch_count(vsr_rt_t* h) { return h->ctx.launches; }
vsr_rt* emu_rt_create() { return new vsr_rt(); }
void emu_rt_destroy(vsr_rt* h) { delete h; }
const char* vsr_last_error(void) { return g_err.c_str(); }
long emu_launches(vsr_rt* h) { return h->ctx.launches; }
}
