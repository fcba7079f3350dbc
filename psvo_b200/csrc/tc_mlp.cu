// tc_mlp.cu -- one-hidden-layer MLP_transformation on sm_100a tensor cores (tcgen05.mma, TMEM accumulators).
//
// Reference: MLP_transformation.transform (src/transformation/MLP.py:48-54) for rows that are particles:
// d_in <= 16, one hidden layer of width H (multiple of 16, <= 256), d_out <= 32 -- the shape of q1 / f / g /
// q1_inv at BASELINE config 5.  Used by psvo_mlp_forward when the shape fits and the row count is large, and
// by the FFBSi step's per-particle nets.  Arithmetic: fp16x3 split products with 2^8 pre-scaled weights, fp32
// accumulation (tc_common.cuh, smc_propagate_tc.cu).
//
// One CTA = 128 threads = one 128-row tile at a time (TMEM lane == row == thread):
//   threads : load the row, split to bf16 hi/lo, store as the A operand                        (CUDA cores)
//   thread 0: layer 1 -- 3 MMAs  D1[128 x H] = A1 * W1^T            (K = 16)                    (tensor core)
//   threads : tcgen05.ld their D1 row, + bias, ReLU, split, store as the A operand of layer 2  (CUDA cores)
//   thread 0: layer 2 -- 3 * H/16 MMAs  D2[128 x NP] = A2 * W2^T    (K = H)                     (tensor core)
//   threads : tcgen05.ld their D2 row, + bias, write mu
#include "smc_kernels.h"
#include "tc_common.cuh"

namespace psvo {

struct TcMlpSmem {
  // byte offsets into dynamic shared memory (all multiples of 128)
  uint32_t w1_hi, w1_lo, w2_hi, w2_lo, a1_hi, a1_lo, a2_hi, a2_lo, b1, b2, bar, slot, total;
};

__host__ __device__ inline TcMlpSmem tc_mlp_layout(int H, int NP) {
  TcMlpSmem s;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
  s.w1_hi = take(H * 16 * 2); s.w1_lo = take(H * 16 * 2);          // [K=16][N=H]
  s.w2_hi = take(NP * H * 2); s.w2_lo = take(NP * H * 2);          // [K=H][N=NP]
  s.a1_hi = take(128 * 16 * 2); s.a1_lo = take(128 * 16 * 2);      // [K=16][128 rows]
  s.a2_hi = take(128 * H * 2); s.a2_lo = take(128 * H * 2);        // [K=H][128 rows]
  s.b1 = take(H * 4); s.b2 = take(NP * 4);
  s.bar = take(16); s.slot = take(16);
  s.total = o;
  return s;
}

template <int NP>
__global__ void __launch_bounds__(128) particle_mlp_tc_kernel(psvo_mlp net, const float* __restrict__ in, long long seq_stride,
                                                              int B, int N, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int H = net.width[0], din = net.d_in, dout = net.d_out;
  const TcMlpSmem L = tc_mlp_layout(H, NP);
  const int tid = threadIdx.x, warp = tid >> 5;
  float* sb1 = reinterpret_cast<float*>(sm + L.b1);
  float* sb2 = reinterpret_cast<float*>(sm + L.b2);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + L.bar);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + L.slot);

  tc::stage_weight_f16x2(net.W[0], din, H, 16, H, 256.0f, sm + L.w1_hi, sm + L.w1_lo);
  tc::stage_weight_f16x2(net.W[1], H, dout, H, NP, 256.0f, sm + L.w2_hi, sm + L.w2_lo);
  for (int i = tid; i < H; i += blockDim.x) sb1[i] = __ldg(net.b[0] + i);
  for (int i = tid; i < NP; i += blockDim.x) sb2[i] = i < dout ? __ldg(net.b[1] + i) : 0.0f;
  if (tid == 0) tc::mbar_init(mbar, 1);
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  const uint32_t tD1 = tmem, tD2 = tmem + 256;                       // column offsets
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc1 = tc::make_idesc_f16(128, H), idesc2 = tc::make_idesc_f16(128, NP);
  const uint32_t a1h = tc::smem_u32(sm + L.a1_hi), a1l = tc::smem_u32(sm + L.a1_lo);
  const uint32_t a2h = tc::smem_u32(sm + L.a2_hi), a2l = tc::smem_u32(sm + L.a2_lo);
  const uint32_t w1h = tc::smem_u32(sm + L.w1_hi), w1l = tc::smem_u32(sm + L.w1_lo);
  const uint32_t w2h = tc::smem_u32(sm + L.w2_hi), w2l = tc::smem_u32(sm + L.w2_lo);
  uint32_t phase = 0;

  const long long rows = (long long)B * N;
  const long long tiles = (rows + 127) / 128;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r = tile * 128 + tid;
    const bool live = r < rows;
    // ---- A1 <- this thread's row
    {
      float x[16];
      const int b = live ? (int)(r / N) : 0, n = live ? (int)(r - (long long)b * N) : 0;
      const float* src = in + b * seq_stride + (long long)n * din;
#pragma unroll
      for (int d = 0; d < 16; ++d) x[d] = (live && d < din) ? src[d] : 0.0f;
#pragma unroll
      for (int kg = 0; kg < 2; ++kg) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = x[kg * 8 + i];
        uint4 hi, lo;
        tc::split8_f16(v, hi, lo);
        *reinterpret_cast<uint4*>(sm + L.a1_hi + kg * 2048 + tid * 16) = hi;
        *reinterpret_cast<uint4*>(sm + L.a1_lo + kg * 2048 + tid * 16) = lo;
      }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    // ---- layer 1
    if (tid < 32 && tc::elect_one()) {
      tc::fence_after_sync();
      const uint64_t dah = tc::make_desc(a1h, 2048, 128), dal = tc::make_desc(a1l, 2048, 128);
      const uint64_t dbh = tc::make_desc(w1h, H * 16, 128), dbl = tc::make_desc(w1l, H * 16, 128);
      tc::mma_bf16(tD1, dah, dbh, idesc1, false);
      tc::mma_bf16(tD1, dah, dbl, idesc1, true);
      tc::mma_bf16(tD1, dal, dbh, idesc1, true);
      tc::mma_commit(mbar);
    }
    tc::mbar_wait(mbar, phase); phase ^= 1;
    tc::fence_after_sync();
    // ---- hidden epilogue: D1 row -> relu(+b1) -> A2 (hi, lo)
    for (int c = 0; c < H; c += 16) {
      float v[16];
      tc::tmem_ld16(tD1 + lane_base + c, v);
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = fmaxf(fmaf(v[g * 8 + i], 1.0f / 256.0f, sb1[c + g * 8 + i]), 0.0f);
        uint4 hi, lo;
        tc::split8_f16(h, hi, lo);
        const int kg = (c >> 3) + g;
        *reinterpret_cast<uint4*>(sm + L.a2_hi + kg * 2048 + tid * 16) = hi;
        *reinterpret_cast<uint4*>(sm + L.a2_lo + kg * 2048 + tid * 16) = lo;
      }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    // ---- layer 2
    if (tid < 32 && tc::elect_one()) {
      tc::fence_after_sync();
      for (int s = 0; s < H / 16; ++s) {
        const uint64_t dah = tc::make_desc(a2h + s * 4096, 2048, 128), dal = tc::make_desc(a2l + s * 4096, 2048, 128);
        const uint64_t dbh = tc::make_desc(w2h + s * (2 * NP * 16), NP * 16, 128), dbl = tc::make_desc(w2l + s * (2 * NP * 16), NP * 16, 128);
        tc::mma_bf16(tD2, dah, dbh, idesc2, s > 0);
        tc::mma_bf16(tD2, dah, dbl, idesc2, true);
        tc::mma_bf16(tD2, dal, dbh, idesc2, true);
      }
      tc::mma_commit(mbar);
    }
    tc::mbar_wait(mbar, phase); phase ^= 1;
    tc::fence_after_sync();
    // ---- output epilogue
#pragma unroll
    for (int c = 0; c < NP; c += 16) {
      float v[16];
      tc::tmem_ld16(tD2 + lane_base + c, v);
      if (live) {
#pragma unroll
        for (int i = 0; i < 16; ++i) if (c + i < dout) out[r * dout + c + i] = fmaf(v[i], 1.0f / 256.0f, sb2[c + i]);
      }
    }
    tc::fence_before_sync();
    __syncthreads();                 // TMEM rows consumed before the next tile's MMAs overwrite them
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

bool tc_mlp_supported(const psvo_mlp& net) {
  return net.n_hidden == 1 && net.d_in >= 1 && net.d_in <= 16 && net.d_out >= 1 && net.d_out <= 32 &&
         net.width[0] >= 16 && net.width[0] <= 256 && (net.width[0] % 16) == 0;
}

// rows of (B sequences x N particles): row (b,n) at in + b*seq_stride + n*d_in ; out contiguous (B*N, d_out)
int launch_particle_mlp_tc(const psvo_mlp& net, const float* in, long long seq_stride, int B, int N, float* out,
                           cudaStream_t st) {
  PSVO_REQUIRE(tc_mlp_supported(net), "tensor-core MLP: unsupported shape");
  const int NP = net.d_out <= 16 ? 16 : 32;
  const TcMlpSmem L = tc_mlp_layout(net.width[0], NP);
  const size_t smem = L.total + 128;
  const long long tiles = ((long long)B * N + 127) / 128;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  ProfScope ps_(PK_ROWS, st);
  if (NP == 16) {
    PSVO_DYN_SMEM(particle_mlp_tc_kernel<16>, 227 * 1024);
    particle_mlp_tc_kernel<16><<<grid, 128, smem, st>>>(net, in, seq_stride, B, N, out);
  } else {
    PSVO_DYN_SMEM(particle_mlp_tc_kernel<32>, 227 * 1024);
    particle_mlp_tc_kernel<32><<<grid, 128, smem, st>>>(net, in, seq_stride, B, N, out);
  }
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo
