// ops.cu -- stand-alone operators of the C ABI (API parity with the reference's helper methods + unit tests).
//   psvo_resample      SVO.get_resample_idx    src/SMC/SVO.py:266-300
//   psvo_gather        SVO.resample_X gather   src/SMC/SVO.py:254-257
//   psvo_get_sigma     tf_mvn.get_sigma        src/distribution/mvn.py:80-90
//   psvo_mvn_log_prob  MultivariateNormalDiag.log_prob   src/distribution/mvn.py:63,104-107
//   psvo_logsumexp     tf.reduce_logsumexp (last axis)
//   psvo_fill_*        the in-kernel Philox streams, materialised
#include "smc_kernels.h"

namespace psvo {

// one CTA per row of logits (K classes), n draws per row
__global__ void __launch_bounds__(1024) resample_rows_kernel(const float* __restrict__ logits, const float* __restrict__ u,
                                                            int K, int n, int* __restrict__ idx) {
  extern __shared__ __align__(16) unsigned char smraw[];
  unsigned long long* sC = reinterpret_cast<unsigned long long*>(smraw);
  float* sL = reinterpret_cast<float*>(sC + K);
  __shared__ unsigned long long sscr[33];
  __shared__ float fscr[33];
  const long long row = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* l = logits + row * K;
  float mx = -INFINITY;
  for (int i = tid; i < K; i += nt) { const float v = l[i]; sL[i] = v; mx = fmaxf(mx, v); }
  mx = block_max(mx, fscr);
  const int ipt = (K + nt - 1) / nt;
  const int lo = min(tid * ipt, K), hi = min(lo + ipt, K);
  unsigned long long run = 0ull;
  for (int i = lo; i < hi; ++i) { run += fixed_weight(sL[i], mx); sC[i] = run; }
  unsigned long long total;
  const unsigned long long off = block_exscan_u64(run, sscr, &total);
  for (int i = lo; i < hi; ++i) sC[i] += off;
  __syncthreads();
  for (int j = tid; j < n; j += nt)
    idx[row * n + j] = upper_bound_u64(sC, K, draw_threshold(total, u24_from_float(u[row * n + j])));
}

__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ idx, long long rows,
                                   int K, int n, int D, float* __restrict__ dst) {
  const long long total = rows * n * (long long)D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long rj = i / D;
    const long long r = rj / n;
    int a = idx[rj];
    a = a < 0 ? 0 : (a >= K ? K - 1 : a);
    dst[i] = src[(r * K + a) * D + d];
  }
}

__global__ void get_sigma_kernel(const float* __restrict__ sc, float smin, int D, float* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) out[d] = sigma_from_con(sc[d], smin);
}

__global__ void mvn_log_prob_kernel(const float* __restrict__ x, const float* __restrict__ mu,
                                    const float* __restrict__ sigma, long long rows, int D, float* __restrict__ out) {
  float ld = 0.0f;
  if (sigma != nullptr) for (int d = 0; d < D; ++d) ld += logf(sigma[d]);
  const float c = 0.5f * (float)D * kLog2Pi;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float z = (x[r * D + d] - mu[r * D + d]) / (sigma != nullptr ? sigma[d] : 1.0f);
      s = fmaf(z, z, s);
    }
    out[r] = -0.5f * s - ld - c;
  }
}

// one warp per row
__global__ void logsumexp_rows_kernel(const float* __restrict__ x, long long rows, int K, float* __restrict__ out) {
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = x + row * K;
  float mx = -INFINITY;
  for (int i = lane; i < K; i += 32) mx = fmaxf(mx, p[i]);
  mx = warp_max(mx);
  const float shift = (mx == -INFINITY || mx == INFINITY || mx != mx) ? 0.0f : mx;   // tf.reduce_logsumexp: non-finite max -> 0
  float s = 0.0f;
  for (int i = lane; i < K; i += 32) s += expf(p[i] - shift);
  s = warp_sum(s);
  if (lane == 0) out[row] = logf(s) + shift;
}

__global__ void fill_normal_kernel(unsigned long long seed, unsigned stream_id, long long first, long long particles,
                                   int width, float* __restrict__ out) {
  const int chunks = (width + 3) / 4;
  const long long total = particles * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / chunks;
    const int c = (int)(i - p * chunks);
    float z[4];
    normal4(philox4x32_10(seed, (uint64_t)(first + p), (uint32_t)c, stream_id), z);
    for (int k = 0; k < 4; ++k) if (4 * c + k < width) out[p * width + 4 * c + k] = z[k];
  }
}

__global__ void fill_uniform_kernel(unsigned long long seed, unsigned stream_id, long long first, long long particles,
                                    float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < particles; i += (long long)gridDim.x * blockDim.x)
    out[i] = uniform_from_bits(uniform_bits(seed, (uint64_t)(first + i), stream_id));
}

static int grid_for(long long n, int block) {
  const long long g = (n + block - 1) / block;
  const long long cap = 32LL * num_sms();
  return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace psvo

using namespace psvo;

extern "C" {

PSVO_API int psvo_resample(const float* logits, const float* u, int64_t rows, int32_t K, int32_t n, int32_t* idx,
                  psvo_stream_t stream) {
  PSVO_REQUIRE(logits && u && idx, "psvo_resample: null pointer");
  PSVO_REQUIRE(rows >= 0 && K >= 1 && n >= 0, "psvo_resample: bad sizes rows=%lld K=%d n=%d", (long long)rows, K, n);
  if (rows == 0 || n == 0) return 0;
  const size_t smem = (size_t)K * (sizeof(unsigned long long) + sizeof(float));
  PSVO_REQUIRE(smem <= 200 * 1024, "psvo_resample: K=%d too large for the shared-memory CDF", K);
  PSVO_DYN_SMEM(resample_rows_kernel, smem);
  const int work = K > n ? K : n;
  const int block = work >= 4096 ? 512 : (work >= 1024 ? 256 : (work >= 256 ? 128 : (work > 32 ? 64 : 32)));
  PSVO_REQUIRE(rows <= 0x7fffffffLL, "psvo_resample: too many rows");
  ProfScope ps_(PK_OPS, (cudaStream_t)stream);
  resample_rows_kernel<<<(unsigned)rows, block, smem, (cudaStream_t)stream>>>(logits, u, K, n, idx);
  PSVO_LAUNCH_CHECK();
  return 0;
}

PSVO_API int psvo_gather(const float* src, const int32_t* idx, int64_t rows, int32_t K, int32_t n, int32_t D, float* dst,
                psvo_stream_t stream) {
  PSVO_REQUIRE(src && idx && dst, "psvo_gather: null pointer");
  PSVO_REQUIRE(rows >= 0 && K >= 1 && n >= 0 && D >= 1, "psvo_gather: bad sizes");
  const long long total = (long long)rows * n * D;
  if (total == 0) return 0;
  ProfScope ps_(PK_OPS, (cudaStream_t)stream);
  gather_rows_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, idx, rows, K, n, D, dst);
  PSVO_LAUNCH_CHECK();
  return 0;
}

PSVO_API int psvo_get_sigma(const float* sigma_con, float sigma_min, int32_t D, float* sigma, psvo_stream_t stream) {
  PSVO_REQUIRE(sigma_con && sigma && D >= 1, "psvo_get_sigma: bad arguments");
  ProfScope ps_(PK_OPS, (cudaStream_t)stream);
  get_sigma_kernel<<<ceil_div(D, 128), 128, 0, (cudaStream_t)stream>>>(sigma_con, sigma_min, D, sigma);
  PSVO_LAUNCH_CHECK();
  return 0;
}

PSVO_API int psvo_mvn_log_prob(const float* x, const float* mu, const float* sigma, int64_t rows, int32_t D, float* out,
                      psvo_stream_t stream) {
  PSVO_REQUIRE(x && mu && out && D >= 1 && rows >= 0, "psvo_mvn_log_prob: bad arguments");
  if (rows == 0) return 0;
  ProfScope ps_(PK_OPS, (cudaStream_t)stream);
  mvn_log_prob_kernel<<<grid_for(rows, 256), 256, 0, (cudaStream_t)stream>>>(x, mu, sigma, rows, D, out);
  PSVO_LAUNCH_CHECK();
  return 0;
}

PSVO_API int psvo_logsumexp(const float* x, int64_t rows, int32_t K, float* out, psvo_stream_t stream) {
  PSVO_REQUIRE(x && out && K >= 1 && rows >= 0, "psvo_logsumexp: bad arguments");
  if (rows == 0) return 0;
  const long long threads = (long long)rows * 32;
  const long long grid = (threads + 255) / 256;
  PSVO_REQUIRE(grid <= 0x7fffffffLL, "psvo_logsumexp: too many rows");
  ProfScope ps_(PK_OPS, (cudaStream_t)stream);
  logsumexp_rows_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(x, rows, K, out);
  PSVO_LAUNCH_CHECK();
  return 0;
}

PSVO_API int psvo_fill_normal(uint64_t seed, int32_t stream_id, int64_t first_particle, int64_t particles, int32_t width,
                     float* out, psvo_stream_t stream) {
  PSVO_REQUIRE(out && particles >= 0 && width >= 1, "psvo_fill_normal: bad arguments");
  if (particles == 0) return 0;
  ProfScope ps_(PK_OPS, (cudaStream_t)stream);
  fill_normal_kernel<<<grid_for(particles * ((width + 3) / 4), 256), 256, 0, (cudaStream_t)stream>>>(
      seed, (unsigned)stream_id, first_particle, particles, width, out);
  PSVO_LAUNCH_CHECK();
  return 0;
}

PSVO_API int psvo_fill_uniform(uint64_t seed, int32_t stream_id, int64_t first_particle, int64_t particles, float* out,
                      psvo_stream_t stream) {
  PSVO_REQUIRE(out && particles >= 0, "psvo_fill_uniform: bad arguments");
  if (particles == 0) return 0;
  ProfScope ps_(PK_OPS, (cudaStream_t)stream);
  fill_uniform_kernel<<<grid_for(particles, 256), 256, 0, (cudaStream_t)stream>>>(seed, (unsigned)stream_id,
                                                                                 first_particle, particles, out);
  PSVO_LAUNCH_CHECK();
  return 0;
}

PSVO_API int psvo_mlp_forward(const psvo_mlp* net, const float* x, int64_t rows, float* mu, psvo_stream_t stream) {
  PSVO_REQUIRE(net && x && mu, "psvo_mlp_forward: null pointer");
  // particle-shaped nets with many rows run on the tensor cores (tc_mlp.cu), everything else on the row kernel
  if (rows >= 256 && rows < (1LL << 31) && tc_mlp_supported(*net))
    return launch_particle_mlp_tc(*net, x, 0, 1, (int)rows, mu, (cudaStream_t)stream);
  return launch_mlp_rows_forward(*net, x, rows, mu, (cudaStream_t)stream);
}

PSVO_API int psvo_mlp_backward(const psvo_mlp* net, const float* x, const float* g_mu, int64_t rows, float* g_x,
                      psvo_mlp_grad* grads, psvo_stream_t stream) {
  PSVO_REQUIRE(net && x && g_mu, "psvo_mlp_backward: null pointer");
  return launch_mlp_rows_backward(*net, x, g_mu, rows, g_x, grads, (cudaStream_t)stream);
}

}  // extern "C"
