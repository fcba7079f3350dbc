// umma_issue_bench.cu -- how many cycles does one small tcgen05.mma (M=128, K=16, fp16, SS mode) cost?
// Sweeps N, the number of issuing threads (1 or 4 warps each issuing its own accumulator chain) and whether
// consecutive MMAs hit the same accumulator.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_issue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../psvo_b200/csrc/tc_common.cuh"
using namespace psvo;

__global__ void __launch_bounds__(160, 1) bench(int N, int n_issuers, int n_acc, int iters, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bars[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // fp16 1.0
  if (tid < 4) tc::mbar_init(bars + tid, 1);
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t idesc = tc::make_idesc_f16(128, N);
  const uint64_t a = tc::make_desc(tc::smem_u32(sm), 2048, 128);                 // A: [K=16][128 rows]
  const uint64_t b = tc::make_desc(tc::smem_u32(sm + 8192), N * 16, 128);        // B: [K=16][N rows]
  long long t0 = 0, t1 = 0;
  if (warp < n_issuers && lane == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + warp * 128 + ((i % n_acc) * N) % 128;             // rotate over n_acc accumulators
      tc::mma_bf16(d, a, b, idesc, i >= n_acc);
    }
    tc::mma_commit(bars + warp);
    tc::mbar_wait(bars + warp, 0);
    t1 = clock64();
    out[blockIdx.x * 4 + warp] = t1 - t0;
  }
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2048;
  printf("%4s %8s %6s %14s\n", "N", "issuers", "n_acc", "cycles/MMA(per issuer)");
  for (int N : {16, 32, 64, 128}) for (int iss : {1, 4}) for (int nacc : {1, 2, 4}) {
    if (nacc * N > 128) continue;
    bench<<<148, 160, 64 * 1024>>>(N, iss, nacc, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("%4d %8d %6d %14.1f\n", N, iss, nacc, (double)h[0] / iters);
  }
  // N = 256 needs all 512 columns for two accumulators: single issuer, one accumulator
  bench<<<148, 160, 64 * 1024>>>(256, 1, 1, iters, d);
  cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("%4d %8d %6d %14.1f\n", 256, 1, 1, (double)h[0] / iters);
  return 0;
}
