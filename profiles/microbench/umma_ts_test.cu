// umma_ts_test.cu -- tcgen05.mma with the A operand in tensor memory (the "TS" form), as needed to keep the hidden
// activations of the particle MLPs out of shared memory:  D[128 x N] (+)= A[128 x K] * B[N x K]^T,
//   A: fp16, written by the row's own thread with tcgen05.st (row r = lane r, column c holds k = 2c, 2c+1),
//      INTO THE COLUMNS OF A PREVIOUS fp32 ACCUMULATOR (aliasing, like the layer-1 accumulator -> layer-2 operand hand-off)
//   B: fp16 K-major no-swizzle in shared memory (element (n, k) at (k/8)*N*16 + n*16 + (k%8)*2)
// Prints max |D - ref|.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../psvo_b200/csrc -o umma_ts umma_ts_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace psvo;
constexpr int K = 32, N = 32;

__global__ void __launch_bounds__(128) probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char* sB = sm;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 8192 + 64);
  const int tid = threadIdx.x;
  for (int i = tid; i < N * K; i += 128) { const int n = i / K, k = i % K; *reinterpret_cast<__half*>(sB + (k / 8) * N * 16 + n * 16 + (k % 8) * 2) = B[n * K + k]; }
  if (tid == 0) tc::mbar_init(bar, 1);
  if (tid < 32) tc::tmem_alloc(slot, 128);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  const uint32_t lane_base = (uint32_t)((tid >> 5) * 32) << 16;
  // A row of this thread: K halves = K/2 columns, at columns [64, 64 + K/2); pad the x32 store with zeros
  uint32_t r[32];
  for (int c = 0; c < 32; ++c) r[c] = c < K / 2 ? *reinterpret_cast<const uint32_t*>(A + (size_t)tid * K + 2 * c) : 0u;
  tc::tmem_st32(tmem + 64 + lane_base, r);
  tc::tmem_st_wait();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid < 32 && tc::elect_one()) {
    const uint32_t idesc = tc::make_idesc_f16(128, N);
    for (int ks = 0; ks < K / 16; ++ks)
      tc::mma_f16_ts(tmem, tmem + 64 + ks * 8, tc::make_desc(tc::smem_u32(sB) + ks * 2 * N * 16, N * 16, 128), idesc, ks > 0);
    tc::mma_commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  float v[32];
  tc::tmem_ld32(tmem + lane_base, v);
  for (int o = 0; o < N; ++o) out[(size_t)tid * N + o] = v[o];
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, 128);
}

int main() {
  __half hA[128 * K], hB[N * K]; float ref[128 * N];
  srand(2);
  for (int i = 0; i < 128 * K; ++i) hA[i] = __float2half((rand() % 17 - 8) / 8.0f);
  for (int i = 0; i < N * K; ++i) hB[i] = __float2half((rand() % 13 - 6) / 4.0f);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += __half2float(hA[m * K + k]) * __half2float(hB[n * K + k]); ref[m * N + n] = s; }
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dO, sizeof(ref));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, sizeof(ref));
  probe<<<1, 128, 16384>>>(dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  float got[128 * N]; cudaMemcpy(got, dO, sizeof(got), cudaMemcpyDeviceToHost);
  float err = 0; for (int i = 0; i < 128 * N; ++i) err = fmaxf(err, fabsf(got[i] - ref[i]));
  printf("TS MMA (A in TMEM, M=128 N=%d K=%d): %s  max |D - ref| = %.5f  (ref[0..3] %.3f %.3f %.3f %.3f got %.3f %.3f %.3f %.3f)\n", N, K,
         cudaGetErrorString(e), err, ref[0], ref[1], ref[2], ref[3], got[0], got[1], got[2], got[3]);
  return err < 1e-3f ? 0 : 1;
}
