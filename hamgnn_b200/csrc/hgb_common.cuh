// Shared helpers for the hamgnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hamgnn_b200.h"

namespace hgb {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void timing_begin(int id, void* stream);
void timing_end(int id, void* stream);
struct TimeScope {   // brackets the kernel launches of one scope with CUDA events when hgb_timing_enable(1) is in effect
  int id;
  void* st;
  TimeScope(int id_, void* st_) : id(id_), st(st_) { timing_begin(id, st); }
  ~TimeScope() { timing_end(id, st); }
};

// Kernels must launch on the device that owns the buffers, whatever the caller's current device is (a model moved to
// cuda:1 without cudaSetDevice(1)): switch to the device of `p` for the duration of the entry point.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const void* p) {
    cudaPointerAttributes at;
    if (p && cudaPointerGetAttributes(&at, p) == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
      if (cudaGetDevice(&prev) == cudaSuccess && prev != at.device) switched = (cudaSetDevice(at.device) == cudaSuccess);
    } else {
      (void)cudaGetLastError();
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define HGB_DEVICE_GUARD(ptr) hgb::DeviceGuard _hgb_device_guard(ptr)

#define HGB_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      hgb::set_error(__VA_ARGS__);      \
      return 1;                         \
    }                                   \
  } while (0)

#define HGB_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      hgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                             \
    }                                                                                       \
  } while (0)

// launch check: configuration errors surface here; execution errors surface at the caller's sync
#define HGB_LAUNCH_OK(name)                                                        \
  do {                                                                             \
    cudaError_t _e = cudaGetLastError();                                           \
    if (_e != cudaSuccess) {                                                       \
      hgb::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
      return 3;                                                                    \
    }                                                                              \
    hgb::count_launch();                                                           \
  } while (0)

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
// softplus(x) - ln 2, torch semantics (beta=1, threshold=20)
__device__ __forceinline__ float ssp_f(float x) {
  float sp = (x > 20.0f) ? x : log1pf(expf(x));
  return sp - 0.69314718055994530942f;
}

__device__ __forceinline__ void fma4x4(float (&acc)[4][4], const float4& a, const float4& w) {
  acc[0][0] = fmaf(a.x, w.x, acc[0][0]); acc[0][1] = fmaf(a.x, w.y, acc[0][1]);
  acc[0][2] = fmaf(a.x, w.z, acc[0][2]); acc[0][3] = fmaf(a.x, w.w, acc[0][3]);
  acc[1][0] = fmaf(a.y, w.x, acc[1][0]); acc[1][1] = fmaf(a.y, w.y, acc[1][1]);
  acc[1][2] = fmaf(a.y, w.z, acc[1][2]); acc[1][3] = fmaf(a.y, w.w, acc[1][3]);
  acc[2][0] = fmaf(a.z, w.x, acc[2][0]); acc[2][1] = fmaf(a.z, w.y, acc[2][1]);
  acc[2][2] = fmaf(a.z, w.z, acc[2][2]); acc[2][3] = fmaf(a.z, w.w, acc[2][3]);
  acc[3][0] = fmaf(a.w, w.x, acc[3][0]); acc[3][1] = fmaf(a.w, w.y, acc[3][1]);
  acc[3][2] = fmaf(a.w, w.z, acc[3][2]); acc[3][3] = fmaf(a.w, w.w, acc[3][3]);
}

}  // namespace hgb
