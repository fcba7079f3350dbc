// umma_throughput_bench.cu -- saturation throughput of small tcgen05.mma (M=128, K=16, fp16, SS) with many issuing warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../psvo_b200/csrc/tc_common.cuh"
using namespace psvo;

__global__ void __launch_bounds__(512, 1) bench(int N, int n_issuers, int iters, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bars[16];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid < 16) tc::mbar_init(bars + tid, 1);
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t idesc = tc::make_idesc_f16(128, N);
  // every issuer reads its own A tile (4 KB) and the shared B tile
  const uint64_t a = tc::make_desc(tc::smem_u32(sm + 16384 + (warp % 8) * 4096), 2048, 128);
  const uint64_t b = tc::make_desc(tc::smem_u32(sm), N * 16, 128);
  __syncthreads();
  const long long t0 = clock64();
  (void)lane;
  if (warp < n_issuers && tc::elect_one()) {      // elect.sync: no per-lane serialisation loop around the MMAs
    for (int i = 0; i < iters; ++i) tc::mma_bf16(tmem + (warp % 2) * 256, a, b, idesc, i > 0);
    tc::mma_commit(bars + warp);
    tc::mbar_wait(bars + warp, 0);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 1024;
  printf("%4s %8s %22s %22s\n", "N", "issuers", "cycles/MMA per issuer", "pipe cycles per MMA");
  for (int N : {16, 32, 64, 128, 256}) for (int iss : {1, 2, 4, 8}) {
    bench<<<148, 512, 64 * 1024>>>(N, iss, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%4d %8d %22.1f %22.1f\n", N, iss, (double)h / iters, (double)h / iters / iss);
  }
  return 0;
}
