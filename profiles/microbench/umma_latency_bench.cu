// umma_latency_bench.cu -- round-trip latencies that bound a fine-grained tcgen05 pipeline:
//   issue k MMAs (M=128,N=32,K=16) -> tcgen05.commit -> mbarrier wait ; st.shared + fence.proxy.async ; tcgen05.ld x32
#include <cstdio>
#include <cuda_runtime.h>
#include "../../psvo_b200/csrc/tc_common.cuh"
using namespace psvo;

__global__ void __launch_bounds__(128, 1) bench(int k, int reps, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  tc::fence_proxy_async(); tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t idesc = tc::make_idesc_f16(128, 32);
  const uint64_t a = tc::make_desc(tc::smem_u32(sm), 2048, 128), b = tc::make_desc(tc::smem_u32(sm + 8192), 32 * 16, 128);
  long long t_rt = 0, t_fence = 0, t_ld = 0, t_bar = 0;
  uint32_t phase = 0;
  if (tid == 0) {
    for (int r = 0; r < reps; ++r) {
      const long long t0 = clock64();
      for (int i = 0; i < k; ++i) tc::mma_bf16(tmem, a, b, idesc, i > 0);
      tc::mma_commit(&bar);
      tc::mbar_wait(&bar, phase); phase ^= 1;
      t_rt += clock64() - t0;
    }
  }
  __syncthreads();
  {  // st.shared + fence.proxy.async
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      *reinterpret_cast<uint4*>(sm + 16384 + tid * 16) = make_uint4(r, r, r, r);
      tc::fence_proxy_async();
    }
    t_fence = clock64() - t0;
  }
  {  // tcgen05.ld x32 + wait
    float v[32]; float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) { tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v); acc += v[r & 31]; }
    t_ld = clock64() - t0;
    if (acc == 123.456f) out[100] = 1;
  }
  {  // named barrier over 128 threads
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) asm volatile("bar.sync 1, 128;" ::: "memory");
    t_bar = clock64() - t0;
  }
  if (tid == 0) { out[0] = t_rt / reps; out[1] = t_fence / reps; out[2] = t_ld / reps; out[3] = t_bar / reps; }
  tc::fence_before_sync(); __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 1024);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int k : {1, 3, 6}) {
    bench<<<1, 128, 64 * 1024>>>(k, 256, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("k=%d MMAs: issue+commit+wait round trip %lld cycles | st.shared+fence.proxy.async %lld | tcgen05.ld.x32+wait %lld | bar.sync(128) %lld\n",
           k, h[0], h[1], h[2], h[3]);
  }
  return 0;
}
