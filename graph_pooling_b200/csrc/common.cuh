// Shared helpers for the gp_b200 kernels (error plumbing, launch accounting, warp reductions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/gp_b200.h"

namespace gp {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;   // process-wide: autograd runs the backward on its own thread

inline int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define GP_REQUIRE(cond, ...) do { if (!(cond)) return gp::fail(GP_ERR_INVALID, __VA_ARGS__); } while (0)
#define GP_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return gp::fail(GP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
// after a <<<>>> launch
#define GP_LAUNCHED() do { gp::g_launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
  return gp::fail(GP_ERR_CUDA, "launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define GP_TRY(call) do { int r_ = (call); if (r_ != 0) return r_; } while (0)

inline cudaStream_t S(gp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum of one float per thread (result valid in every thread); `sh` >= 33 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? sh[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// (graph, node) of a flattened row index r = b * N + n.  Every caller has B * N < 2^31 rows (checked on the host), so the
// division runs in 32 bits: a 64-bit division by a run-time divisor costs ~60 instructions per row and lane, as much as
// the useful work of a 128-float row.
__device__ __forceinline__ void row_split(long long r, int N, int& b, int& n) {
  const unsigned ur = (unsigned)r, un = (unsigned)N;
  const unsigned ub = ur / un;
  b = (int)ub;
  n = (int)(ur - ub * un);
}

constexpr int kNumSMs = 148;   // B200 (grid sizing only: a device with another SM count runs the same kernels correctly)

// "Configure once per DEVICE": cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device / context, so the
// done-flag is one bit per device ordinal.  The attribute is set BEFORE the bit is published: a concurrent first call
// from another thread at worst sets it twice (harmless), and a process that moves to a second GPU configures it there.
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool pending(unsigned long long* bit) {
    int d = 0;
    cudaGetDevice(&d);
    *bit = 1ull << (d & 63);
    return !(done.load(std::memory_order_acquire) & *bit);
  }
  void mark(unsigned long long bit) { done.fetch_or(bit, std::memory_order_release); }
};
#define GP_CONFIG_ONCE(...)                                   \
  do {                                                        \
    static gp::DeviceOnce once_;                              \
    unsigned long long bit_;                                  \
    if (once_.pending(&bit_)) { __VA_ARGS__; once_.mark(bit_); } \
  } while (0)

}  // namespace gp
