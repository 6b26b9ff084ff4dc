// Shared helpers of libhgsfa: error reporting, dtype sizes, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "hgsfa.h"

namespace hgsfa {

std::string& last_error_ref();
int fail(const char* fmt, ...);

#define HG_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ::hgsfa::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                         \
  } while (0)

#define HG_CHECK(cond, ...)                         \
  do {                                              \
    if (!(cond)) return ::hgsfa::fail(__VA_ARGS__); \
  } while (0)

inline size_t dtype_size(int dt) { return dt == HGSFA_U8 ? 1 : (dt == HGSFA_F32 ? 4 : 8); }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// RAII device guard: every entry point runs on its handle's device and restores the caller's.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Entry points that take raw device buffers and no handle run on the device that owns the buffer, whatever
// device is current in the calling thread (a detector bound to cuda:k is driven from threads whose current
// device is 0).  Falls back to the current device for pointers CUDA does not know.
inline int device_of(const void* p) {
  int cur = 0;
  cudaGetDevice(&cur);
  if (!p) return cur;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return cur;
  }
  return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : cur;
}
struct PtrDeviceGuard : DeviceGuard {
  int device;
  explicit PtrDeviceGuard(const void* p) : PtrDeviceGuard(device_of(p), 0) {}
 private:
  PtrDeviceGuard(int dev, int) : DeviceGuard(dev), device(dev) {}
};

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need) {
    if (need <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) return fail("cudaMalloc(%zu bytes) failed: %s", need, cudaGetErrorString(e));
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
};

}  // namespace hgsfa
