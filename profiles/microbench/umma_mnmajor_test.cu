// umma_mnmajor_test.cu -- probe for the reverse-mode kernels: weight-gradient GEMMs  dW[j][o] = sum_p H[p][j] * G[p][o]
// on tcgen05 with BOTH operands MN-major, i.e. read straight from the row-wise buffers the MLP epilogue writes
// (thread p stores 8 consecutive j as one 16-byte chunk: element (p, j) at (j/8)*P*16 + p*16 + (j%8)*2).
//   A = H^T : M = j (64 or 128), K = p (128)      B = G^T : N = o (32), K = p
// Canonical MN-major no-swizzle layout (CUTLASS mma_sm100_desc): ((T,1,m),(8,k)):((1,T,SBO),(1T,LBO)), T = 8 halves:
//   element (mn, k) at (mn%8)*2 + (k%8)*16 + (k/8)*LBO + (mn/8)*SBO   ->  LBO = 128 B, SBO = P*16 B for our buffers.
// Prints max |D - ref| and where the M rows of the accumulator land in TMEM (lane map) for M = 64 and 128.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I ../../psvo_b200/csrc -o umma_mn umma_mnmajor_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"

using namespace psvo;

constexpr int P = 128;      // particles = K
constexpr int NO = 32;      // outputs = N

__global__ void __launch_bounds__(128) probe(const __half* __restrict__ H, const __half* __restrict__ G, float* __restrict__ out,
                                             int M, uint32_t lbo, uint32_t sbo) {
  extern __shared__ __align__(128) unsigned char sm[];
  unsigned char* sH = sm;                       // [M/8][P][8] halves
  unsigned char* sG = sm + 32768;               // [NO/8][P][8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 49152);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 49152 + 64);
  const int tid = threadIdx.x;
  // row-wise staging exactly like the MLP epilogue: thread p writes its 16-byte chunks
  for (int jg = 0; jg < M / 8; ++jg) *reinterpret_cast<uint4*>(sH + jg * P * 16 + tid * 16) = *reinterpret_cast<const uint4*>(H + (size_t)tid * M + jg * 8);
  for (int og = 0; og < NO / 8; ++og) *reinterpret_cast<uint4*>(sG + og * P * 16 + tid * 16) = *reinterpret_cast<const uint4*>(G + (size_t)tid * NO + og * 8);
  if (tid == 0) tc::mbar_init(bar, 1);
  if (tid < 32) tc::tmem_alloc(slot, 64);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_f16(M, NO) | (1u << 15) | (1u << 16);     // a_major = b_major = MN
    for (int ks = 0; ks < P / 16; ++ks) {
      const uint64_t a = tc::make_desc(tc::smem_u32(sH) + ks * 2 * 128, lbo, sbo);   // 16 k = two k-groups of 8 (128 B each)
      const uint64_t b = tc::make_desc(tc::smem_u32(sG) + ks * 2 * 128, lbo, sbo);
      tc::mma_bf16(tmem, a, b, idesc, ks > 0);
    }
    tc::mma_commit(bar);
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  float v[32];
  tc::tmem_ld32(tmem + ((uint32_t)((tid >> 5) * 32) << 16), v);
  for (int o = 0; o < NO; ++o) out[(size_t)tid * NO + o] = v[o];               // lane tid of TMEM, all 128 lanes
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 32) tc::tmem_dealloc(tmem, 64);
}

int main() {
  for (int M : {128, 64}) {
    __half *hH = new __half[P * M], *hG = new __half[P * NO];
    float* ref = new float[M * NO]();
    srand(1);
    for (int i = 0; i < P * M; ++i) hH[i] = __float2half((rand() % 17 - 8) / 8.0f);
    for (int i = 0; i < P * NO; ++i) hG[i] = __float2half((rand() % 13 - 6) / 4.0f);
    for (int j = 0; j < M; ++j) for (int o = 0; o < NO; ++o) { float s = 0; for (int p = 0; p < P; ++p) s += __half2float(hH[p * M + j]) * __half2float(hG[p * NO + o]); ref[j * NO + o] = s; }
    __half *dH, *dG; float* dO;
    cudaMalloc(&dH, P * M * 2); cudaMalloc(&dG, P * NO * 2); cudaMalloc(&dO, 128 * NO * 4);
    cudaMemcpy(dH, hH, P * M * 2, cudaMemcpyHostToDevice); cudaMemcpy(dG, hG, P * NO * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const uint32_t cand[2][2] = {{128, (uint32_t)P * 16}, {(uint32_t)P * 16, 128}};   // (LBO, SBO) as derived, and swapped
    for (int c = 0; c < 2; ++c) {
      cudaMemset(dO, 0, 128 * NO * 4);
      probe<<<1, 128, 65536>>>(dH, dG, dO, M, cand[c][0], cand[c][1]);
      cudaError_t e = cudaDeviceSynchronize();
      float got[128 * NO];
      cudaMemcpy(got, dO, sizeof(got), cudaMemcpyDeviceToHost);
      // identity lane map
      float err_id = 0; for (int j = 0; j < M; ++j) for (int o = 0; o < NO; ++o) err_id = fmaxf(err_id, fabsf(got[j * NO + o] - ref[j * NO + o]));
      // search: for each row j find the lane whose 32 values match
      int lane_of[128]; int found = 0;
      for (int j = 0; j < M; ++j) { lane_of[j] = -1; for (int l = 0; l < 128; ++l) { float er = 0; for (int o = 0; o < NO; ++o) er = fmaxf(er, fabsf(got[l * NO + o] - ref[j * NO + o])); if (er < 1e-3f) { lane_of[j] = l; ++found; break; } } }
      printf("M=%d LBO=%u SBO=%u: %s  max err (row j = lane j) %.4f ; rows located %d/%d ; lanes of rows 0,1,15,16,31,32,63: %d %d %d %d %d %d %d\n", M,
             cand[c][0], cand[c][1], cudaGetErrorString(e), err_id, found, M, lane_of[0], lane_of[1], lane_of[15], lane_of[16], lane_of[31], lane_of[32], lane_of[63]);
    }
  }
  return 0;
}
