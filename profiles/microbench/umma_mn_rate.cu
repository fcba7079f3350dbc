// umma_mn_rate.cu -- cost of the MMA shapes used by mlp_bwd_tc.cu (fp16, K = 16 per instruction), one issuing thread,
// 512 back-to-back MMAs into one accumulator, clock64 from first issue to commit completion.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../psvo_b200/csrc -o umma_mn_rate umma_mn_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace psvo;

__global__ void __launch_bounds__(128) rate(long long* out, int M, int N, int a_mn, int b_mn, uint32_t lbo, uint32_t sbo, int reps) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 98304);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 98304 + 64);
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) tc::mbar_init(bar, 1);
  if (threadIdx.x < 32) tc::tmem_alloc(slot, 256);
  tc::fence_proxy_async(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32 && tc::elect_one()) {      // elect.sync: no per-lane serialisation loop around the MMAs
    const uint32_t idesc = tc::make_idesc_f16(M, N) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
    const uint64_t a = tc::make_desc(tc::smem_u32(sm), lbo, sbo), b = tc::make_desc(tc::smem_u32(sm) + 49152, lbo, sbo);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) tc::mma_bf16(tmem, a, b, idesc, i > 0);
    tc::mma_commit(bar);
    tc::mbar_wait(bar, 0);
    out[blockIdx.x] = clock64() - t0;
  }
  tc::fence_before_sync(); __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 256);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000);
  struct C { const char* name; int M, N, a, b; uint32_t lbo, sbo; } cs[] = {
      {"K-major  M128 N64  (L1, dh)          ", 128, 64, 0, 0, 2048, 128},
      {"K-major  M128 N16  (dx, fwd L2)      ", 128, 16, 0, 0, 2048, 128},
      {"MN-major M64  N32  SBO 2048 (dW2)    ", 64, 32, 1, 1, 128, 2048},
      {"MN-major M64  N32  SBO 2064 (padded) ", 64, 32, 1, 1, 128, 2064},
      {"MN-major M64  N24  SBO 2064 (dW1)    ", 64, 24, 1, 1, 128, 2064},
      {"MN-major M128 N32  SBO 2064          ", 128, 32, 1, 1, 128, 2064},
      {"A K-major, B MN-major M128 N16 (dx)  ", 128, 16, 0, 1, 2064, 128},
  };
  for (auto& c : cs) {
    rate<<<148, 128, 100000>>>(d, c.M, c.N, c.a, c.b, c.lbo, c.sbo, 512);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s %s  %.1f cycles per MMA (one elected issuer, 148 SMs busy)\n", c.name, cudaGetErrorString(e), mx / 512.0);
  }
  return 0;
}
