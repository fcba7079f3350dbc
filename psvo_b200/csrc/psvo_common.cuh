// psvo_common.cuh -- host-side error plumbing and block-level device primitives.
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/psvo.h"
#include "psvo_math.cuh"

namespace psvo {

// ---- error plumbing (include/psvo.h: 0 ok, <0 argument error with text, >0 cudaError_t) ----------------
std::string& last_error();
int fail(const char* fmt, ...);
#define PSVO_REQUIRE(cond, ...) do { if (!(cond)) return ::psvo::fail(__VA_ARGS__); } while (0)
#define PSVO_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
    ::psvo::fail("%s failed: %s", #expr, cudaGetErrorString(e_)); return (int)e_; } } while (0)
#define PSVO_LAUNCH_CHECK() PSVO_CUDA(cudaGetLastError())

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// pad a dimension to the next instantiated register-array size
inline int pad_dx(int d) { return d <= 4 ? 4 : (d <= 8 ? 8 : 16); }
inline int pad_dy(int d) { return d <= 4 ? 4 : (d <= 8 ? 8 : (d <= 16 ? 16 : 32)); }

int num_sms();
// Raise a kernel's dynamic shared-memory limit; remembered per (device, kernel) because the attribute is per device: one process
// may drive several GPUs.  Returns 0 or the cudaError_t (with psvo_last_error() set).
int ensure_dyn_smem(const void* kernel, size_t bytes);
#define PSVO_DYN_SMEM(kern, bytes) do { if (int e_ = ::psvo::ensure_dyn_smem(reinterpret_cast<const void*>(kern), (size_t)(bytes))) return e_; } while (0)

// ---- measurement hooks (include/psvo.h: psvo_profile_*) ------------------------------------------------
enum { PK_PROP = 0, PK_RESAMPLE, PK_ROWS, PK_SIGMA, PK_BWD, PK_FFBSI, PK_OPS, PK_SPARE };
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
struct ProfScope {
  int kind; cudaStream_t st;
  ProfScope(int k, cudaStream_t s) : kind(k), st(s) { prof_begin(k, s); }
  ~ProfScope() { prof_end(kind, st); }
};

#if defined(__CUDACC__)
// ---- warp / block primitives ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// scratch: >= 33 floats of shared memory.  All threads of the block must call.  Result broadcast to all.
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? scratch[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
// exclusive block scan of one uint64 per thread; returns the exclusive prefix, *total = block total.
// scratch: >= 33 uint64.  Integer addition is associative, so the result does not depend on the scan shape.
__device__ __forceinline__ unsigned long long block_exscan_u64(unsigned long long v, unsigned long long* scratch,
                                                               unsigned long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = lane < nw ? scratch[lane] : 0ull;
    unsigned long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    if (lane < nw) scratch[lane] = winc - w;          // exclusive warp offsets
    if (lane == 31) scratch[32] = winc;               // grand total (lanes >= nw contribute 0)
  }
  __syncthreads();
  *total = scratch[32];
  return scratch[warp] + inc - v;
}
#endif  // __CUDACC__

// Programmatic dependent launch: the kernel may start (prologue: shared-memory staging, TMEM allocation, barrier init) while the
// previous kernel on the stream is still draining; it must execute griddep_wait() before touching anything that kernel wrote.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
// PSVO_PDL=0 switches programmatic dependent launches off (plain stream order everywhere)
inline bool pdl_enabled() {
  static const bool on = !(getenv("PSVO_PDL") != nullptr && getenv("PSVO_PDL")[0] == '0');
  return on;
}
#if defined(__CUDACC__)
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// entry of a small kernel inside a launch-latency-bound loop: let the next launch be scheduled, then wait for the previous kernel
__device__ __forceinline__ void griddep_entry() { griddep_launch_dependents(); griddep_wait(); }
#endif

}  // namespace psvo
