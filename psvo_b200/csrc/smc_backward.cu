// smc_backward.cu -- reverse mode of the forward particle filter.
//
// Reference: what optimizer.minimize(-log_ZSMC) (src/trainer.py:115-118) derives through SVO.SMC
// (src/SMC/SVO.py:60-180): path-wise gradients through the reparameterised samples X_t = m + s*eps and through
// the ancestor gather; tf.multinomial has no gradient, so nothing flows into the resampling logits
// (SURVEY.md 3.5, App. A.8).  With resampling, log Wbar_{t-1} = -log N is a constant, so
//     d logZ_b / d log_W_t[n] = softmax_n(log_W_t)[n].
// Without resampling (IWAE) the per-step log-sum-exps telescope and d logZ_b / d log_alpha_t[n] is the FINAL
// softmax weight of particle n for every t.
//
// Time runs backwards; per step two kernels (the MLP hidden activations are recomputed, never stored):
//   bwd_emission : per particle, MLP g forward + reverse   -> gXg[b][n][:] = d/dX_t via the emission,
//                  accumulates dW_g and the sigma_g statistic.
//   bwd_proposal : per particle, MLP q1 (and f) forward + reverse, fusion / sample / f,q densities reverse
//                  -> d/dX_anc_{t-1}, scattered (atomicAdd) onto the parents X_{t-1}[anc]; accumulates dW_q1,
//                  dW_f, the per-(b,t) gradient of the q2 mean and the sigma statistics.
// Afterwards the row-MLP reverse mode back-propagates the q2 / q0 / f(pre_x0) means to their parameters and
// to pre_obs / pre_x0, and a tiny kernel maps the sigma statistics to d/d sigma_con.
#include "mlp_bwd_tile.cuh"
#include <algorithm>
#include "smc_kernels.h"

namespace psvo {

// statistic slots accumulated by the step kernels (each 32 floats): index = slot*2 + tau  (tau = 0: t = 0)
enum { ST_SP = 0, ST_I1, ST_I2, ST_SF, ST_SG, ST_SLOTS };

struct BwdArgs {
  psvo_mlp q1, f, g;
  psvo_mlp_grad gq1, gf, gg;
  const float* sig;
  const float* obs; const float* mu2_all; const float* mu0; const float* muf0;
  const float* Xs; const float* X_anc; const int* anc; const float* log_Ws; const float* lse;
  const float* g_logZ_b;
  const float* ext_mu;    // tensor-core path: the network output of every particle (B*N, d_out) computed by a tcgen05
  float* ext_g;           //   kernel; this kernel then only writes the upstream gradient (B*N, d_out) for mlp_bwd_tc
  const float* g_Xs;      // (B,T,N,Dx) incoming d/dXs (FFBSi reverse mode) or NULL
  const float* g_log_Ws;  // (B,T,N)    incoming d/dlog_Ws or NULL
  float* gXg;        // (B,N,Dx) emission contribution to d/dX_t         (written by bwd_emission)
  const float* gXin; // (B,N,Dx) children contribution to d/dX_t         (scattered by step t+1)
  float* gXout;      // (B,N,Dx) d/dX_{t-1} accumulator (zeroed before the launch)
  float* gmu2_all;   // (B,T,Dx)
  float* gmu0;       // (B,Dx)
  float* gmuf0;      // (B,Dx)
  float* stats;      // [ST_SLOTS*2][32]
  int t_all = 0;                  // bwd_emission only: 1 = ALL time steps in this launch (gXg is then (T,B,N,Dx)); it depends on forward outputs only
  unsigned int* gmax = nullptr;   // tensor-core path: max |ext_g| of the launch (bit pattern; mlp_bwd_tc's operand scale), or NULL
  int N, B, T, Dx, Dy, t;
  unsigned flags;
};

__device__ __forceinline__ float particle_weight_grad_at(const BwdArgs& a, int b, int n, int t, bool live) {
  // d L / d log_W_t[n]  (see file header)
  if (!live) return 0.0f;
  const bool resample = a.flags & PSVO_F_RESAMPLE;
  const size_t bt = (size_t)b * a.T + (resample ? t : a.T - 1);
  float g = expf(a.log_Ws[bt * a.N + n] - a.lse[bt]) * a.g_logZ_b[b];
  if (a.g_log_Ws != nullptr) g += a.g_log_Ws[((size_t)b * a.T + t) * a.N + n];     // resampling mode only (checked by the ABI)
  return g;
}
__device__ __forceinline__ float particle_weight_grad(const BwdArgs& a, int b, int n, bool live) {
  return particle_weight_grad_at(a, b, n, a.t, live);
}

template <int DXP, int DYP>
__global__ void __launch_bounds__(256) bwd_emission_kernel(BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  griddep_entry();
  const int Dx = a.Dx, Dy = a.Dy, N = a.N;
  const bool poisson = a.flags & PSVO_F_POISSON;
  MlpS<DXP, DYP> gs;
  MlpGradS<DXP, DYP> gg;
  float* p = smem;
  float* sg = p; p += 32;
  float* sstat = p; p += 32;
  p = mlp_stage<DXP, DYP>(a.g, p, gs);
  float* gacc0 = p;
  p = mlp_grad_carve<DXP, DYP>(a.g, p, gg);
  for (float* q = gacc0 + threadIdx.x; q < p; q += blockDim.x) *q = 0.0f;
  float* stage = p;
  if (threadIdx.x < 32) { sg[threadIdx.x] = a.sig[SIG_G * 32 + threadIdx.x]; sstat[threadIdx.x] = 0.0f; }
  __syncthreads();

  const int tile = blockDim.x;
  const int tiles_per_seq = (N + tile - 1) / tile;
  const int work_t = a.B * tiles_per_seq;
  const int work = a.t_all ? work_t * a.T : work_t;          // t_all: every (t, b, tile) of the pass in one launch
  float st_sg[DYP];
#pragma unroll
  for (int d = 0; d < DYP; ++d) st_sg[d] = 0.0f;

  for (int wj = blockIdx.x; wj < work; wj += gridDim.x) {
    const int t = a.t_all ? wj / work_t : a.t;
    const int wi = a.t_all ? wj - t * work_t : wj;
    const int b = wi / tiles_per_seq;
    const int n = (wi - b * tiles_per_seq) * tile + threadIdx.x;
    const bool live = n < N;
    const size_t bt = (size_t)b * a.T + t;
    float x[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) x[d] = (live && d < Dx) ? a.Xs[(bt * N + n) * Dx + d] : 0.0f;
    float xx[1][DXP], mug[1][DYP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) xx[0][d] = x[d];
    mlp_eval<DXP, DYP, 1>(gs, xx, mug, nullptr, 0);
    const float gl = particle_weight_grad_at(a, b, n, t, live);
    float G[DYP];
#pragma unroll
    for (int d = 0; d < DYP; ++d) {
      G[d] = 0.0f;
      if (d < Dy) {
        const float y = a.obs[bt * Dy + d];
        if (poisson) {
          const float z = y - (softplus(mug[0][d]) + 1e-6f);
          G[d] = gl * z * sigmoidf(mug[0][d]);
        } else {
          const float z = (y - mug[0][d]) / sg[d];
          G[d] = gl * z / sg[d];
          st_sg[d] += gl * (z * z - 1.0f) / sg[d];
        }
      }
    }
    float gx[DXP];
    mlp_bwd_tile<DXP, DYP>(gs, gg, x, G, gx, stage);
    if (live) {
#pragma unroll
      for (int d = 0; d < DXP; ++d) if (d < Dx) a.gXg[(((size_t)(a.t_all ? t : 0) * a.B + b) * N + n) * Dx + d] = gx[d];
    }
  }
  __syncthreads();
  if (!poisson) {
#pragma unroll
    for (int d = 0; d < DYP; ++d) {
      const float v = warp_sum(st_sg[d]);
      if ((threadIdx.x & 31) == 0 && d < Dy) atomicAdd(&sstat[d], v);
    }
  }
  __syncthreads();
  if (!poisson && threadIdx.x < Dy) atomicAdd(a.stats + (ST_SG * 2 + 1) * 32 + threadIdx.x, sstat[threadIdx.x]);
  mlp_grad_flush<DXP, DYP>(a.g, gg, a.gg);
}

template <int DXP>
__global__ void __launch_bounds__(256) bwd_proposal_kernel(BwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  griddep_entry();
  const int Dx = a.Dx, N = a.N;
  const bool two_q = a.flags & PSVO_F_TWO_Q, boot = a.flags & PSVO_F_BOOTSTRAP;
  const bool t0 = a.t == 0;
  const bool f_is_d1 = boot && two_q, sep_f = !boot;
  const int tau = t0 ? 0 : 1;

  MlpS<DXP, DXP> q1s, fs;
  MlpGradS<DXP, DXP> gq1, gf;
  float* p = smem;
  float* C = p; p += 5 * DXP;                 // s1 | sf | sp | inv1 | inv2
  float* sred = p; p += 4 * 32;               // block partials: gm2 | gmu1 (t0) | gmuf (t0) | spare
  float* sstat = p; p += 4 * 32;              // SP | I1 | I2 | SF
  float* gacc0 = p;
  if (!t0) {
    p = mlp_stage<DXP, DXP>(a.q1, p, q1s);
    if (sep_f) p = mlp_stage<DXP, DXP>(a.f, p, fs);
    gacc0 = p;
    p = mlp_grad_carve<DXP, DXP>(a.q1, p, gq1);
    if (sep_f) p = mlp_grad_carve<DXP, DXP>(a.f, p, gf);
    for (float* q = gacc0 + threadIdx.x; q < p; q += blockDim.x) *q = 0.0f;
  }
  float* stage = p;
  float *s1 = C, *sf = C + DXP, *sp = C + 2 * DXP, *inv1 = C + 3 * DXP, *inv2 = C + 4 * DXP;
  if (threadIdx.x < DXP) {
    const int d = threadIdx.x;
    const float v1 = a.sig[(t0 ? SIG_Q0 : SIG_Q1) * 32 + d], v2 = a.sig[SIG_Q2 * 32 + d];
    s1[d] = v1;
    sf[d] = f_is_d1 ? v1 : a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
    inv1[d] = 1.0f / v1; inv2[d] = 1.0f / v2;
    sp[d] = two_q ? 1.0f / (1.0f / v1 + 1.0f / v2) : v1;
  }
  for (int i = threadIdx.x; i < 4 * 32; i += blockDim.x) { sstat[i] = 0.0f; sred[i] = 0.0f; }
  __syncthreads();

  const int tile = blockDim.x;
  const int tiles_per_seq = (N + tile - 1) / tile;
  const int work = a.B * tiles_per_seq;
  float st_sp[DXP], st_i1[DXP], st_i2[DXP], st_sf[DXP];
#pragma unroll
  for (int d = 0; d < DXP; ++d) st_sp[d] = st_i1[d] = st_i2[d] = st_sf[d] = 0.0f;

  for (int wi = blockIdx.x; wi < work; wi += gridDim.x) {
    const int b = wi / tiles_per_seq;
    const int n = (wi - b * tiles_per_seq) * tile + threadIdx.x;
    const bool live = n < N;
    const size_t bt = (size_t)b * a.T + a.t;
    const size_t bn = (size_t)b * N + (live ? n : 0);

    float par[1][DXP], mu1[1][DXP], muf[1][DXP];
    if (!t0) {
#pragma unroll
      for (int d = 0; d < DXP; ++d) par[0][d] = (live && d < Dx) ? a.X_anc[((bt - 1) * N + n) * Dx + d] : 0.0f;
      mlp_eval<DXP, DXP, 1>(q1s, par, mu1, nullptr, 0);
      if (sep_f) mlp_eval<DXP, DXP, 1>(fs, par, muf, nullptr, 0);
    } else {
#pragma unroll
      for (int d = 0; d < DXP; ++d) {
        par[0][d] = 0.0f;
        mu1[0][d] = d < Dx ? a.mu0[(size_t)b * Dx + d] : 0.0f;
        if (!f_is_d1) muf[0][d] = d < Dx ? a.muf0[(size_t)b * Dx + d] : 0.0f;
      }
    }
    if (f_is_d1 || (!t0 && !sep_f)) {
#pragma unroll
      for (int d = 0; d < DXP; ++d) muf[0][d] = mu1[0][d];
    }
    const float gl = particle_weight_grad(a, b, n, live);

    float gmu1[DXP], gmuf[DXP], gm2[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) {
      gmu1[d] = gmuf[d] = gm2[d] = 0.0f;
      if (d < Dx && live) {
        const float X = a.Xs[(bt * N + n) * Dx + d];
        const float m2 = two_q ? a.mu2_all[bt * Dx + d] : 0.0f;
        const float A = inv1[d] * mu1[0][d] + inv2[d] * m2;              // m / sp  (two_q)
        const float m = two_q ? sp[d] * A : mu1[0][d];
        const float e = (X - m) / sp[d];
        const float zf = (X - muf[0][d]) / sf[d];
        float gX = a.gXin[bn * Dx + d] + a.gXg[bn * Dx + d] - gl * zf / sf[d];
        if (a.g_Xs != nullptr) gX += a.g_Xs[(bt * N + n) * Dx + d];
        // X = m + sp*e ; -q_lp contributes +gl/sp to d/dsp, nothing to d/dm
        st_sp[d] += gX * e + gl / sp[d] + (two_q ? gX * A : 0.0f);
        st_sf[d] += gl * (zf * zf - 1.0f) / sf[d];
        const float gmf = gl * zf / sf[d];                                // d f_lp / d muf
        float g1 = two_q ? gX * sp[d] * inv1[d] : gX;                     // via m
        if (two_q) {
          st_i1[d] += gX * sp[d] * mu1[0][d];
          st_i2[d] += gX * sp[d] * m2;
          gm2[d] = gX * sp[d] * inv2[d];
        }
        if (f_is_d1 || (!t0 && !sep_f)) g1 += gmf;                        // muf IS mu1
        else gmuf[d] = gmf;
        gmu1[d] = g1;
      }
    }

    // per-(b,t) gradient of the q2 mean; at t = 0 also of the per-sequence q0 / f means
    if (two_q || t0) {
      __syncthreads();
#pragma unroll
      for (int d = 0; d < DXP; ++d) {
        if (d < Dx) {
          if (two_q) { const float v = warp_sum(gm2[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sred[d], v); }
          if (t0) {
            const float v1 = warp_sum(gmu1[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sred[32 + d], v1);
            if (!f_is_d1) { const float v2 = warp_sum(gmuf[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sred[64 + d], v2); }
          }
        }
      }
      __syncthreads();
      if (threadIdx.x < Dx) {
        const int d = threadIdx.x;
        if (two_q) { atomicAdd(a.gmu2_all + bt * Dx + d, sred[d]); sred[d] = 0.0f; }
        if (t0) {
          atomicAdd(a.gmu0 + (size_t)b * Dx + d, sred[32 + d]); sred[32 + d] = 0.0f;
          if (!f_is_d1) { atomicAdd(a.gmuf0 + (size_t)b * Dx + d, sred[64 + d]); sred[64 + d] = 0.0f; }
        }
      }
    }

    if (!t0) {
      float ga[DXP], ga2[DXP];
      mlp_bwd_tile<DXP, DXP>(q1s, gq1, par[0], gmu1, ga, stage);
      if (sep_f) {
        mlp_bwd_tile<DXP, DXP>(fs, gf, par[0], gmuf, ga2, stage);
#pragma unroll
        for (int d = 0; d < DXP; ++d) ga[d] += ga2[d];
      }
      if (live) {
        const int parent = a.anc[(bt - 1) * N + n];
        float* dst = a.gXout + ((size_t)b * N + parent) * Dx;
#pragma unroll
        for (int d = 0; d < DXP; ++d) if (d < Dx) atomicAdd(dst + d, ga[d]);
      }
    }
  }

  __syncthreads();
#pragma unroll
  for (int d = 0; d < DXP; ++d) {
    if (d < Dx) {
      float v = warp_sum(st_sp[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[d], v);
      v = warp_sum(st_i1[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[32 + d], v);
      v = warp_sum(st_i2[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[64 + d], v);
      v = warp_sum(st_sf[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[96 + d], v);
    }
  }
  __syncthreads();
  if (threadIdx.x < Dx) {
    const int d = threadIdx.x;
    atomicAdd(a.stats + (ST_SP * 2 + tau) * 32 + d, sstat[d]);
    atomicAdd(a.stats + (ST_I1 * 2 + tau) * 32 + d, sstat[32 + d]);
    atomicAdd(a.stats + (ST_I2 * 2 + tau) * 32 + d, sstat[64 + d]);
    atomicAdd(a.stats + (ST_SF * 2 + tau) * 32 + d, sstat[96 + d]);
  }
  if (!t0) {
    mlp_grad_flush<DXP, DXP>(a.q1, gq1, a.gq1);
    if (sep_f) mlp_grad_flush<DXP, DXP>(a.f, gf, a.gf);
  }
}

// ------------------------------------------------------------------------------------------------------
// Tensor-core path (default flags: use_bootstrap and use_2_q): the MLP forward (mu) and reverse (gx, dW, db) run in
// tcgen05 kernels; these two light kernels do what is left of a step per particle -- the upstream gradient of the network
// output, the sigma statistics and the per-sequence reductions -- with the same formulas as the fused kernels above.
// ------------------------------------------------------------------------------------------------------
// one thread per FOUR consecutive columns of a row (Gaussian emission, Dy a multiple of 4; 128-bit loads and stores, the
// particle's weight gradient computed once per four elements, 32-bit index arithmetic, reciprocal of sigma), else one thread per
// element; fully coalesced reads of mu and writes of G either way
__global__ void __launch_bounds__(256) emission_grad_kernel(BwdArgs a) {
  __shared__ float sstat[32];
  __shared__ float ssg[32], sinv[32];
  const int Dy = a.Dy, N = a.N;
  const bool poisson = a.flags & PSVO_F_POISSON;
  if (threadIdx.x < 32) {
    const float v = a.sig[SIG_G * 32 + threadIdx.x];
    sstat[threadIdx.x] = 0.0f; ssg[threadIdx.x] = v; sinv[threadIdx.x] = 1.0f / v;
  }
  __syncthreads();
  const long long total = (long long)a.B * N * Dy;
  if (!poisson && (Dy & 3) == 0 && total < (1LL << 31) && (256 % (Dy >> 2)) == 0) {
    const int q4 = Dy >> 2;                                  // column groups per row
    const int groups = (int)(total >> 2);
    const int cg = (int)(threadIdx.x % q4);                  // the same column group in every iteration (256 and the stride are multiples of q4)
    const float4 isg = make_float4(sinv[4 * cg], sinv[4 * cg + 1], sinv[4 * cg + 2], sinv[4 * cg + 3]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float gabs = 0.0f;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < groups; e += gridDim.x * 256) {
      const int row = e / q4;
      const int b = row / N, n = row - b * N;
      const size_t bt = (size_t)b * a.T + a.t;
      const float gl = particle_weight_grad(a, b, n, true);
      const float4 y = __ldg(reinterpret_cast<const float4*>(a.obs + bt * Dy) + cg);
      const float4 mu = reinterpret_cast<const float4*>(a.ext_mu)[e];
      float4 g;
      const float z0 = (y.x - mu.x) * isg.x, z1 = (y.y - mu.y) * isg.y, z2 = (y.z - mu.z) * isg.z, z3 = (y.w - mu.w) * isg.w;
      g.x = gl * z0 * isg.x; g.y = gl * z1 * isg.y; g.z = gl * z2 * isg.z; g.w = gl * z3 * isg.w;
      reinterpret_cast<float4*>(a.ext_g)[e] = g;
      gabs = fmaxf(fmaxf(gabs, fmaxf(fabsf(g.x), fabsf(g.y))), fmaxf(fabsf(g.z), fabsf(g.w)));
      acc.x += gl * (z0 * z0 - 1.0f) * isg.x; acc.y += gl * (z1 * z1 - 1.0f) * isg.y;
      acc.z += gl * (z2 * z2 - 1.0f) * isg.z; acc.w += gl * (z3 * z3 - 1.0f) * isg.w;
    }
    if (a.gmax != nullptr) {                                 // the reverse GEMM kernel's operand scale: no pass of its own over G
      gabs = warp_max(gabs);
      if ((threadIdx.x & 31) == 0 && gabs > 0.0f) atomicMax(a.gmax, __float_as_uint(gabs));
    }
    atomicAdd(&sstat[4 * cg], acc.x); atomicAdd(&sstat[4 * cg + 1], acc.y);
    atomicAdd(&sstat[4 * cg + 2], acc.z); atomicAdd(&sstat[4 * cg + 3], acc.w);
    __syncthreads();
    if (threadIdx.x < Dy) { const float v = sstat[threadIdx.x]; if (v != 0.0f) atomicAdd(a.stats + (ST_SG * 2 + 1) * 32 + threadIdx.x, v); }
    return;
  }
  const bool fixed_d = (blockDim.x % Dy) == 0;        // then e % Dy is the same in every iteration of a thread
  float acc = 0.0f, gabs = 0.0f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / Dy;
    const int d = (int)(e - row * Dy);
    const int b = (int)(row / N), n = (int)(row - (long long)b * N);
    const size_t bt = (size_t)b * a.T + a.t;
    const float gl = particle_weight_grad(a, b, n, true);
    const float y = __ldg(a.obs + bt * Dy + d);
    const float mu = a.ext_mu[e];
    if (poisson) {
      const float z = y - (softplus(mu) + 1e-6f);
      const float gv = gl * z * sigmoidf(mu);
      a.ext_g[e] = gv; gabs = fmaxf(gabs, fabsf(gv));
    } else {
      const float sg = ssg[d];
      const float z = (y - mu) / sg;
      const float gv = gl * z / sg;
      a.ext_g[e] = gv; gabs = fmaxf(gabs, fabsf(gv));
      const float st = gl * (z * z - 1.0f) / sg;
      if (fixed_d) acc += st; else atomicAdd(&sstat[d], st);
    }
  }
  if (a.gmax != nullptr) {
    gabs = warp_max(gabs);
    if ((threadIdx.x & 31) == 0 && gabs > 0.0f) atomicMax(a.gmax, __float_as_uint(gabs));
  }
  if (!poisson && fixed_d && acc != 0.0f) atomicAdd(&sstat[threadIdx.x % Dy], acc);
  __syncthreads();
  if (!poisson && threadIdx.x < Dy) { const float v = sstat[threadIdx.x]; if (v != 0.0f) atomicAdd(a.stats + (ST_SG * 2 + 1) * 32 + threadIdx.x, v); }
}

// t >= 1, f aliases d1 = q1:  mu1 = MLP q1(X_anc[t-1]) from ext_mu; writes d/dmu1 to ext_g
__global__ void __launch_bounds__(256, 1) proposal_grad_kernel(BwdArgs a) {
  __shared__ float sstat[4 * 32];
  __shared__ float sp[PSVO_MAX_DX], inv1[PSVO_MAX_DX], inv2[PSVO_MAX_DX];
  const int Dx = a.Dx, N = a.N;
  for (int i = threadIdx.x; i < 4 * 32; i += blockDim.x) sstat[i] = 0.0f;
  if (threadIdx.x < PSVO_MAX_DX) {
    const int d = threadIdx.x;
    const float v1 = d < Dx ? __ldg(a.sig + SIG_Q1 * 32 + d) : 1.0f, v2 = d < Dx ? __ldg(a.sig + SIG_Q2 * 32 + d) : 1.0f;
    inv1[d] = 1.0f / v1; inv2[d] = 1.0f / v2; sp[d] = 1.0f / (1.0f / v1 + 1.0f / v2);
  }
  __syncthreads();
  float st_sp[PSVO_MAX_DX], st_i1[PSVO_MAX_DX], st_i2[PSVO_MAX_DX], st_sf[PSVO_MAX_DX];
#pragma unroll
  for (int d = 0; d < PSVO_MAX_DX; ++d) st_sp[d] = st_i1[d] = st_i2[d] = st_sf[d] = 0.0f;
  float gabs = 0.0f;
  const long long rows = (long long)a.B * N;
  const long long rows_pad = (rows + 31) / 32 * 32;                       // whole warps iterate together (shuffles below)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < rows_pad; idx += (long long)gridDim.x * blockDim.x) {
    const bool live = idx < rows;
    const long long id = live ? idx : rows - 1;
    const int b = (int)(id / N), n = (int)(id - (long long)b * N);
    const size_t bt = (size_t)b * a.T + a.t;
    const float gl = live ? particle_weight_grad(a, b, n, true) : 0.0f;
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const bool uniform = __all_sync(0xffffffffu, b == b0);                // the warp's particles belong to one sequence
    // the row's four inputs in registers first (128-bit loads when Dx is a multiple of 4): one latency instead of 4*Dx
    float r_mu[PSVO_MAX_DX], r_x[PSVO_MAX_DX], r_gi[PSVO_MAX_DX], r_gg[PSVO_MAX_DX];
    {
      const float* p_mu = a.ext_mu + id * Dx; const float* p_x = a.Xs + (bt * N + n) * Dx;
      const float* p_gi = a.gXin + id * Dx;   const float* p_gg = a.gXg + id * Dx;
      if ((Dx & 3) == 0) {
#pragma unroll
        for (int q = 0; q < PSVO_MAX_DX / 4; ++q) {
          if (4 * q < Dx) {
            const float4 v0 = *reinterpret_cast<const float4*>(p_mu + 4 * q), v1 = *reinterpret_cast<const float4*>(p_x + 4 * q);
            const float4 v2 = *reinterpret_cast<const float4*>(p_gi + 4 * q), v3 = *reinterpret_cast<const float4*>(p_gg + 4 * q);
            r_mu[4 * q] = v0.x; r_mu[4 * q + 1] = v0.y; r_mu[4 * q + 2] = v0.z; r_mu[4 * q + 3] = v0.w;
            r_x[4 * q] = v1.x; r_x[4 * q + 1] = v1.y; r_x[4 * q + 2] = v1.z; r_x[4 * q + 3] = v1.w;
            r_gi[4 * q] = v2.x; r_gi[4 * q + 1] = v2.y; r_gi[4 * q + 2] = v2.z; r_gi[4 * q + 3] = v2.w;
            r_gg[4 * q] = v3.x; r_gg[4 * q + 1] = v3.y; r_gg[4 * q + 2] = v3.z; r_gg[4 * q + 3] = v3.w;
          }
        }
      } else {
#pragma unroll
        for (int d = 0; d < PSVO_MAX_DX; ++d) if (d < Dx) { r_mu[d] = p_mu[d]; r_x[d] = p_x[d]; r_gi[d] = p_gi[d]; r_gg[d] = p_gg[d]; }
      }
    }
    float gm2v[PSVO_MAX_DX], gout[PSVO_MAX_DX], r_m2[PSVO_MAX_DX];
    {
      const float* p_m2 = a.mu2_all + bt * Dx;                           // the sequence's second proposal mean: one 128-bit row read
      if ((Dx & 3) == 0) {
#pragma unroll
        for (int q = 0; q < PSVO_MAX_DX / 4; ++q) {
          if (4 * q < Dx) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p_m2) + q);
            r_m2[4 * q] = v.x; r_m2[4 * q + 1] = v.y; r_m2[4 * q + 2] = v.z; r_m2[4 * q + 3] = v.w;
          }
        }
      } else {
#pragma unroll
        for (int d = 0; d < PSVO_MAX_DX; ++d) if (d < Dx) r_m2[d] = __ldg(p_m2 + d);
      }
    }
#pragma unroll
    for (int d = 0; d < PSVO_MAX_DX; ++d) {
      gm2v[d] = 0.0f; gout[d] = 0.0f;
      if (d < Dx) {
        float gm2 = 0.0f;
        if (live) {
          const float mu1 = r_mu[d];
          const float X = r_x[d];
          const float m2 = r_m2[d];
          const float A = inv1[d] * mu1 + inv2[d] * m2;                   // m / sp
          const float m = sp[d] * A;
          const float isp = inv1[d] + inv2[d];                            // 1 / sp
          const float e = (X - m) * isp;
          const float zf = (X - mu1) * inv1[d];                           // f_lp = d1_mvn.log_prob(X): sf = s1, muf = mu1
          float gX = r_gi[d] + r_gg[d] - gl * zf * inv1[d];
          if (a.g_Xs != nullptr) gX += a.g_Xs[(bt * N + n) * Dx + d];
          st_sp[d] += gX * e + gl * isp + gX * A;
          st_sf[d] += gl * (zf * zf - 1.0f) * inv1[d];
          st_i1[d] += gX * sp[d] * mu1;
          st_i2[d] += gX * sp[d] * m2;
          gm2 = gX * sp[d] * inv2[d];
          const float gv = gX * sp[d] * inv1[d] + gl * zf * inv1[d];        // via m, plus d f_lp / d muf (muf IS mu1)
          gout[d] = gv; gabs = fmaxf(gabs, fabsf(gv));
        }
        if (uniform) gm2v[d] = gm2;
        else if (live && gm2 != 0.0f) atomicAdd(a.gmu2_all + bt * Dx + d, gm2);
      }
    }
    if (live) {                                                            // the row of upstream gradients: 128-bit stores
      float* p_g = a.ext_g + id * Dx;
      if ((Dx & 3) == 0) {
#pragma unroll
        for (int q = 0; q < PSVO_MAX_DX / 4; ++q)
          if (4 * q < Dx) reinterpret_cast<float4*>(p_g)[q] = make_float4(gout[4 * q], gout[4 * q + 1], gout[4 * q + 2], gout[4 * q + 3]);
      } else {
#pragma unroll
        for (int d = 0; d < PSVO_MAX_DX; ++d) if (d < Dx) p_g[d] = gout[d];
      }
    }
    if (uniform) {
      // the warp's 16 column sums with 16 shuffles instead of 16 x 5: every butterfly step halves the values a lane carries
      // (lane bit 4 picks the half, bit 3 the quarter, ...); lanes 2k and 2k+1 end up with the total of column k's bit pattern
      const int lane = threadIdx.x & 31;
      float w8[8], w4[4], w2[2];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool hi = lane & 16;
        const float keep = hi ? gm2v[i + 8] : gm2v[i], send = hi ? gm2v[i] : gm2v[i + 8];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool hi = lane & 8;
        const float keep = hi ? w8[i + 4] : w8[i], send = hi ? w8[i] : w8[i + 4];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const bool hi = lane & 4;
        const float keep = hi ? w4[i + 2] : w4[i], send = hi ? w4[i] : w4[i + 2];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      const bool hi = lane & 2;
      float v = (hi ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, hi ? w2[0] : w2[1], 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      const size_t bt0 = (size_t)b0 * a.T + a.t;
      if ((lane & 1) == 0 && col < Dx && v != 0.0f) atomicAdd(a.gmu2_all + bt0 * Dx + col, v);
    }
  }
  if (a.gmax != nullptr) {
    gabs = warp_max(gabs);
    if ((threadIdx.x & 31) == 0 && gabs > 0.0f) atomicMax(a.gmax, __float_as_uint(gabs));
  }
#pragma unroll
  for (int d = 0; d < PSVO_MAX_DX; ++d) {
    if (d < Dx) {
      float v = warp_sum(st_sp[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[d], v);
      v = warp_sum(st_i1[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[32 + d], v);
      v = warp_sum(st_i2[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[64 + d], v);
      v = warp_sum(st_sf[d]); if ((threadIdx.x & 31) == 0) atomicAdd(&sstat[96 + d], v);
    }
  }
  __syncthreads();
  if (threadIdx.x < Dx) {
    const int d = threadIdx.x;
    atomicAdd(a.stats + (ST_SP * 2 + 1) * 32 + d, sstat[d]);
    atomicAdd(a.stats + (ST_I1 * 2 + 1) * 32 + d, sstat[32 + d]);
    atomicAdd(a.stats + (ST_I2 * 2 + 1) * 32 + d, sstat[64 + d]);
    atomicAdd(a.stats + (ST_SF * 2 + 1) * 32 + d, sstat[96 + d]);
  }
}

// statistics -> d/d sigma_con  (one block, 32 threads)
struct SigmaGradArgs {
  const float* stats; const float* sig;
  const float* con[SIG_SLOTS]; float smin[SIG_SLOTS]; float* gcon[SIG_SLOTS];
  int Dx, Dy; unsigned flags;
};
__global__ void sigma_grad_kernel(SigmaGradArgs a) {
  const int d = threadIdx.x;
  const bool two_q = a.flags & PSVO_F_TWO_Q, boot = a.flags & PSVO_F_BOOTSTRAP, poisson = a.flags & PSVO_F_POISSON;
  const bool f_is_d1 = boot && two_q;
  float gs[SIG_SLOTS];
  for (int k = 0; k < SIG_SLOTS; ++k) gs[k] = 0.0f;
  if (d < a.Dx) {
    for (int tau = 0; tau < 2; ++tau) {
      const int d1 = tau == 0 ? SIG_Q0 : SIG_Q1;
      const float s1 = a.sig[d1 * 32 + d], s2 = a.sig[SIG_Q2 * 32 + d];
      const float acc_sp = a.stats[(ST_SP * 2 + tau) * 32 + d];
      if (two_q) {
        const float i1 = 1.0f / s1, i2 = 1.0f / s2, sp = 1.0f / (i1 + i2);
        const float gi1 = a.stats[(ST_I1 * 2 + tau) * 32 + d] - sp * sp * acc_sp;
        const float gi2 = a.stats[(ST_I2 * 2 + tau) * 32 + d] - sp * sp * acc_sp;
        gs[d1] += -gi1 * i1 * i1;
        gs[SIG_Q2] += -gi2 * i2 * i2;
      } else {
        gs[d1] += acc_sp;
      }
      const int owner = f_is_d1 ? d1 : (boot ? SIG_Q1 : SIG_F);
      gs[owner] += a.stats[(ST_SF * 2 + tau) * 32 + d];
    }
  }
  if (!poisson && d < a.Dy) gs[SIG_G] += a.stats[(ST_SG * 2 + 1) * 32 + d];
  for (int k = 0; k < SIG_SLOTS; ++k) {
    const int dim = k == SIG_G ? a.Dy : a.Dx;
    if (a.gcon[k] == nullptr || a.con[k] == nullptr || d >= dim) continue;
    const float sc = a.con[k][d];
    const float sval = softplus(sc);
    const bool pass = (sc == sc) && (sval > a.smin[k]);       // NaN scrub and clamp have zero gradient
    if (pass) atomicAdd(a.gcon[k] + d, gs[k] * sigmoidf(sc));
  }
}

template <int DXP, int DYP>
static int launch_emission_t(const BwdArgs& a, int block, int grid, size_t smem, cudaStream_t st) {
  auto kern = bwd_emission_kernel<DXP, DYP>;
  PSVO_DYN_SMEM(kern, smem);
  ProfScope ps_(PK_BWD, st);
  PSVO_CUDA(launch_pdl(kern, dim3(grid), dim3(block), smem, st, pdl_enabled(), a));
  PSVO_LAUNCH_CHECK();
  return 0;
}
template <int DXP>
static int launch_proposal_t(const BwdArgs& a, int block, int grid, size_t smem, cudaStream_t st) {
  auto kern = bwd_proposal_kernel<DXP>;
  PSVO_DYN_SMEM(kern, smem);
  ProfScope ps_(PK_BWD, st);
  PSVO_CUDA(launch_pdl(kern, dim3(grid), dim3(block), smem, st, pdl_enabled(), a));
  PSVO_LAUNCH_CHECK();
  return 0;
}

static size_t staged_floats(const psvo_mlp& net, int din_pad, int dout_pad) {
  if (net.n_hidden == 0) return (size_t)din_pad * dout_pad + dout_pad;
  return (size_t)net.width[0] * din_pad + round4(net.width[0]) + (size_t)net.width[0] * dout_pad + dout_pad;
}

struct BwdWs {
  float *sig, *mu2_all, *mu0, *muf0, *gmu2_all, *gmu0, *gmuf0, *gXa, *gXb, *gXg, *stats, *gx0tmp;
  float* gXg_all;      // (T,B,N,Dx): emission contributions of all steps from ONE launch (small passes on the FP32 kernels), or NULL
  float *tc_mu, *tc_muq, *tc_g; unsigned char *img_g, *img_q1, *img_fwd; unsigned int* gmax;     // tensor-core reverse mode (or NULL)
  float *rows_x, *rows_gx;                                                                       // generic row-MLP reverse mode: deep networks (or NULL)
  size_t bytes;
};
static bool tc_backward_ok(const psvo_dims* d, const psvo_params* p) {
  if (p == nullptr) return false;
  if (!(d->flags & PSVO_F_BOOTSTRAP) || !(d->flags & PSVO_F_TWO_Q)) return false;      // f aliases d1: one network per step side
  const long long rows = (long long)d->B * d->N;
  return mlp_bwd_tc_supported(p->q1.net, rows) && mlp_bwd_tc_supported(p->g.net, rows) &&
         tc_pipeline_supported(p->q1.net, p->g.net, d->Dx, d->Dy) && rows < (1LL << 31);
}
// per-particle networks with more than one hidden layer: reverse mode through the generic row-MLP kernels (any depth), with
// this file's upstream-gradient kernels in between -- the structure of the tensor-core path, default flag combination only
static bool deep_rows_backward(const psvo_dims* d, const psvo_params* p) {
  if (p == nullptr) return false;
  if (!(d->flags & PSVO_F_BOOTSTRAP) || !(d->flags & PSVO_F_TWO_Q)) return false;
  return p->q1.net.n_hidden >= 2 || p->g.net.n_hidden >= 2;
}
static BwdWs carve_bwd(const psvo_dims* d, const psvo_params* p, char* base) {
  BwdWs w;
  size_t off = 0;
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(base + off); off += align_up(floats * sizeof(float), 256); return r; };
  const size_t BT = (size_t)d->B * d->T, BN = (size_t)d->B * d->N;
  w.sig = take(SIG_SLOTS * 32);
  w.mu2_all = take(BT * d->Dx); w.mu0 = take((size_t)d->B * d->Dx); w.muf0 = take((size_t)d->B * d->Dx);
  w.gmu2_all = take(BT * d->Dx); w.gmu0 = take((size_t)d->B * d->Dx); w.gmuf0 = take((size_t)d->B * d->Dx);
  w.gXa = take(BN * d->Dx); w.gXb = take(BN * d->Dx); w.gXg = take(BN * d->Dx);
  w.gXg_all = nullptr;
  if (p != nullptr && !tc_backward_ok(d, p) && !deep_rows_backward(d, p) && (size_t)d->T * BN * d->Dx <= (size_t)1 << 22)
    w.gXg_all = take((size_t)d->T * BN * d->Dx);
  w.stats = take(ST_SLOTS * 2 * 32);
  w.gx0tmp = take((size_t)d->B * d->Dx0);
  w.tc_mu = w.tc_muq = w.tc_g = nullptr; w.img_g = w.img_q1 = w.img_fwd = nullptr; w.gmax = nullptr;
  if (tc_backward_ok(d, p)) {
    const size_t wide = d->Dx > d->Dy ? d->Dx : d->Dy;
    w.tc_mu = take(BN * d->Dy); w.tc_muq = take(BN * d->Dx); w.tc_g = take(BN * wide);
    w.img_fwd = reinterpret_cast<unsigned char*>(take(32 * 1024));
    w.img_g = reinterpret_cast<unsigned char*>(take(mlp_bwd_tc_image_bytes(p->g.net) / sizeof(float) + 64));
    w.img_q1 = reinterpret_cast<unsigned char*>(take(mlp_bwd_tc_image_bytes(p->q1.net) / sizeof(float) + 64));
    w.gmax = reinterpret_cast<unsigned int*>(take(64));
  }
  w.rows_x = w.rows_gx = nullptr;
  if (!tc_backward_ok(d, p) && deep_rows_backward(d, p)) {
    const size_t wide = d->Dx > d->Dy ? d->Dx : d->Dy;
    w.tc_mu = take(BN * d->Dy); w.tc_muq = take(BN * d->Dx); w.tc_g = take(BN * wide);
    w.rows_x = take(BN * d->Dx); w.rows_gx = take(BN * d->Dx);
  }
  w.bytes = off;
  return w;
}

// gXout[b][anc[b][n]][:] += gx[b][n][:]  (a child's gradient flows to its resampled parent, SVO.py:254-257 under autodiff)
__global__ void scatter_parent_kernel(const float* __restrict__ gx, const int* __restrict__ anc, long long anc_seq_stride,
                                      float* __restrict__ gXout, int B, int N, int Dx) {
  const long long total = (long long)B * N * Dx;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / Dx;
    const int d = (int)(e - row * Dx);
    const int b = (int)(row / N), n = (int)(row - (long long)b * N);
    const int parent = anc[(long long)b * anc_seq_stride + n];
    atomicAdd(gXout + ((long long)b * N + parent) * Dx + d, gx[e]);
  }
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}

int check_dims(const psvo_dims* d);
int check_smc_params(const psvo_dims* d, const psvo_params* p);
SigmaSrc sigma_src(const psvo_dims* d, const psvo_params* p, bool ffbsi);

}  // namespace psvo

using namespace psvo;

extern "C" {

PSVO_API size_t psvo_smc_backward_workspace_bytes(const psvo_dims* d, const psvo_params* p) {
  if (d == nullptr) return 0;
  return carve_bwd(d, p, nullptr).bytes;
}

PSVO_API int psvo_smc_backward(const psvo_dims* d, const psvo_params* p, const float* obs, const float* pre_obs,
                               const float* pre_x0, const float* eps, uint64_t seed, const float* Xs, const float* X_anc,
                               const int32_t* anc, const float* log_Ws, const float* lse, const float* g_logZ_b,
                               const float* g_Xs, const float* g_log_Ws, psvo_grads* grads, float* g_pre_obs, float* g_pre_x0, void* ws, size_t ws_bytes,
                               psvo_stream_t stream) {
  (void)eps; (void)seed;     // the noise is recovered as (X - m)/s from the stored particles
  if (int e = check_dims(d)) return e;
  if (int e = check_smc_params(d, p)) return e;
  PSVO_REQUIRE(obs && pre_x0 && Xs && X_anc && anc && log_Ws && lse && g_logZ_b && grads && ws, "psvo_smc_backward: null buffer");
  const bool two_q = d->flags & PSVO_F_TWO_Q, boot = d->flags & PSVO_F_BOOTSTRAP;
  const bool f_is_d1 = boot && two_q, sep_f = !boot;
  PSVO_REQUIRE(!two_q || pre_obs, "psvo_smc_backward: pre_obs is null but use_2_q is set");
  PSVO_REQUIRE(g_log_Ws == nullptr || (d->flags & PSVO_F_RESAMPLE),
               "psvo_smc_backward: an incoming g_log_Ws is only supported with resampling (PSVO / PSVOwR always resample)");
  const psvo_dist& fd = boot ? p->q1 : p->f;
  const bool deep = p->q1.net.n_hidden >= 2 || p->g.net.n_hidden >= 2 || fd.net.n_hidden >= 2;
  PSVO_REQUIRE(!deep || f_is_d1,
               "psvo_smc_backward: per-particle networks (q1, f, g) with more than one hidden layer need use_bootstrap and use_2_q in reverse mode");
  const BwdWs w = carve_bwd(d, p, static_cast<char*>(ws));
  PSVO_REQUIRE(ws_bytes >= w.bytes, "workspace too small: %zu < %zu", ws_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, T = d->T, N = d->N, Dx = d->Dx;
  const size_t BT = (size_t)B * T, BN = (size_t)B * N;

  // recompute the per-sequence means, clear the accumulators
  if (int e = launch_sigma_block(sigma_src(d, p, false), w.sig, st)) return e;
  if (two_q) if (int e = launch_mlp_rows_forward(p->q2.net, pre_obs, (int64_t)BT, w.mu2_all, st)) return e;
  if (int e = launch_mlp_rows_forward(p->q0.net, pre_x0, B, w.mu0, st)) return e;
  if (!f_is_d1) if (int e = launch_mlp_rows_forward(fd.net, pre_x0, B, w.muf0, st)) return e;
  PSVO_CUDA(cudaMemsetAsync(w.gmu2_all, 0, BT * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.gmu0, 0, (size_t)B * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.gmuf0, 0, (size_t)B * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.stats, 0, ST_SLOTS * 2 * 32 * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.gXa, 0, BN * Dx * sizeof(float), st));

  BwdArgs a;
  a.q1 = p->q1.net; a.f = fd.net; a.g = p->g.net;
  a.gq1 = grads->q1.net; a.gf = boot ? grads->q1.net : grads->f.net; a.gg = grads->g.net;
  a.sig = w.sig; a.obs = obs; a.mu2_all = w.mu2_all; a.mu0 = w.mu0; a.muf0 = w.muf0;
  a.Xs = Xs; a.X_anc = X_anc; a.anc = anc; a.log_Ws = log_Ws; a.lse = lse; a.g_logZ_b = g_logZ_b;
  a.g_Xs = g_Xs; a.g_log_Ws = g_log_Ws;
  a.gXg = w.gXg; a.gmu2_all = w.gmu2_all; a.gmu0 = w.gmu0; a.gmuf0 = w.gmuf0; a.stats = w.stats;
  a.N = N; a.B = B; a.T = T; a.Dx = Dx; a.Dy = d->Dy; a.flags = d->flags;

  const int DXP = pad_dx(Dx), DYP = pad_dy(d->Dy);
  // block size: as many particles per tile as shared memory allows (weights + gradient accumulator + staging)
  auto pick_block = [&](size_t fixed_floats, int din, int dout, int* block, size_t* smem) -> int {
    int P = 256;
    while (P > 32 && N <= P / 2) P /= 2;
    for (;; P /= 2) {
      const size_t fl = fixed_floats + (size_t)2 * kBwdChunk * (P + 1) + (size_t)P * din + (size_t)P * dout;
      if (fl * sizeof(float) <= 220 * 1024 || P == 32) { *block = P; *smem = fl * sizeof(float); break; }
    }
    PSVO_REQUIRE(*smem <= 227 * 1024, "psvo_smc_backward: networks too large for shared memory (%zu B)", *smem);
    return 0;
  };
  int blkE, blkP; size_t smE, smP;
  if (int e = pick_block(64 + 2 * staged_floats(a.g, DXP, DYP), DXP, DYP, &blkE, &smE)) return e;
  const size_t fixedP = 5 * DXP + 256 + 2 * staged_floats(a.q1, DXP, DXP) + (sep_f ? 2 * staged_floats(a.f, DXP, DXP) : 0);
  if (int e = pick_block(fixedP, DXP, DXP, &blkP, &smP)) return e;
  // persistent CTAs: as many per SM as shared memory and the thread budget allow (each CTA flushes its gradient
  // accumulator once, so more than a few per SM only adds atomics)
  auto grid_of = [&](int block, size_t smem) {
    const int64_t work = (int64_t)B * ((N + block - 1) / block);
    int per_sm = (int)((220 * 1024) / (smem > 0 ? smem : 1));
    if (per_sm > 2048 / block) per_sm = 2048 / block;
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    const int64_t cap = (int64_t)num_sms() * per_sm;
    return (int)(work < cap ? work : cap);
  };

  // tensor-core reverse mode (default flags, one hidden layer of 64..256 units, >= 8192 particles per step): the two
  // per-particle networks run as  tcgen05 forward (mu) -> this file's kernels without their MLP part (upstream gradient,
  // sigma statistics, per-sequence reductions) -> tcgen05 reverse (mlp_bwd_tc.cu: gx, dW, db)
  const bool tc = w.img_g != nullptr;
  if (tc) {
    PSVO_CUDA(cudaMemsetAsync(w.img_g, 0, mlp_bwd_tc_image_bytes(a.g), st));
    PSVO_CUDA(cudaMemsetAsync(w.img_q1, 0, mlp_bwd_tc_image_bytes(a.q1), st));
    if (int e = launch_mlp_bwd_tc_stage(a.g, w.img_g, st)) return e;
    if (int e = launch_mlp_bwd_tc_stage(a.q1, w.img_q1, st)) return e;
    PropArgs ia;
    ia.q1 = a.q1; ia.f = a.q1; ia.g = a.g; ia.Dx = Dx; ia.Dy = d->Dy;
    if (int e = launch_tc_stage_image(ia, w.img_fwd, st)) return e;
  }
  a.ext_mu = nullptr; a.ext_g = nullptr;
  const int light_grid = (int)(((long long)B * N + 255) / 256 < 8LL * num_sms() ? ((long long)B * N + 255) / 256 : 8LL * num_sms());
  // the proposal kernel ends with 64 warp reductions per thread (sigma statistics): two CTAs per SM, several rows per thread
  const int prop_grid = light_grid < num_sms() ? light_grid : num_sms();

  float* gXin = w.gXa;      // children contribution for the current step (zero at t = T-1)
  float* gXout = w.gXb;
  for (int t = T - 1; t >= 0; --t) {
    a.t = t; a.gXin = gXin; a.gXout = gXout;
    if (t > 0) PSVO_CUDA(cudaMemsetAsync(gXout, 0, BN * Dx * sizeof(float), st));
    if (tc) {
      const long long sstr = (long long)T * N * Dx;
      const float* Xt = Xs + (size_t)t * N * Dx;
      int rc = 0;
      // ---- emission side
      {
        Fwd2TcArgs fa;
        fa.tc_image = w.img_fwd; fa.xq = t > 0 ? X_anc + (size_t)(t - 1) * N * Dx : nullptr; fa.xg = Xt; fa.x_seq_stride = sstr;
        fa.muq = w.tc_muq; fa.mug = w.tc_mu; fa.B = B; fa.N = N; fa.Dx = Dx; fa.Dy = d->Dy; fa.H = a.g.width[0];
        if ((rc = launch_tc_fwd2(fa, st))) return rc;
      }
      // max |G| of both upstream gradients (mlp_bwd_tc's operand scales) comes out of the kernels that write G
      PSVO_CUDA(cudaMemsetAsync(w.gmax, 0, 2 * sizeof(unsigned int), st));
      a.ext_mu = w.tc_mu; a.ext_g = w.tc_g; a.gmax = w.gmax;
      {
        ProfScope ps_(PK_BWD, st);
        emission_grad_kernel<<<light_grid, 256, 0, st>>>(a);
        PSVO_LAUNCH_CHECK();
      }
      if ((rc = launch_mlp_bwd_tc(a.g, w.img_g, Xt, sstr, w.tc_g, B, N, w.gXg, nullptr, 0, a.gg, w.gmax, st, true))) return rc;
      // ---- proposal side
      if (t > 0) {
        const float* Xp = X_anc + (size_t)(t - 1) * N * Dx;
        a.ext_mu = w.tc_muq; a.gmax = w.gmax + 1;
        {
          ProfScope ps_(PK_BWD, st);
          proposal_grad_kernel<<<prop_grid, 256, 0, st>>>(a);
          PSVO_LAUNCH_CHECK();
        }
        if ((rc = launch_mlp_bwd_tc(a.q1, w.img_q1, Xp, sstr, w.tc_g, B, N, gXout, anc + (size_t)(t - 1) * N, (long long)T * N,
                                    a.gq1, w.gmax + 1, st, true))) return rc;
      } else {
        a.ext_mu = nullptr; a.ext_g = nullptr; a.gmax = nullptr;   // t = 0: no per-particle proposal network (q0 is per sequence)
        rc = DXP == 4 ? launch_proposal_t<4>(a, blkP, grid_of(blkP, smP), smP, st)
                      : (DXP == 8 ? launch_proposal_t<8>(a, blkP, grid_of(blkP, smP), smP, st)
                                  : launch_proposal_t<16>(a, blkP, grid_of(blkP, smP), smP, st));
        if (rc) return rc;
      }
      float* tmp = gXin; gXin = gXout; gXout = tmp;
      continue;
    }
    if (w.rows_x != nullptr) {
      // ---- deep networks: forward and reverse through the generic row-MLP kernels on this step's rows, staged contiguously
      const size_t pitch = (size_t)T * N * Dx * sizeof(float), width = (size_t)N * Dx * sizeof(float);
      PSVO_CUDA(cudaMemcpy2DAsync(w.rows_x, width, Xs + (size_t)t * N * Dx, pitch, width, B, cudaMemcpyDeviceToDevice, st));
      if (int e = launch_mlp_rows_forward(a.g, w.rows_x, (int64_t)BN, w.tc_mu, st)) return e;
      a.ext_mu = w.tc_mu; a.ext_g = w.tc_g; a.gmax = nullptr;
      {
        ProfScope ps_(PK_BWD, st);
        emission_grad_kernel<<<light_grid, 256, 0, st>>>(a);
        PSVO_LAUNCH_CHECK();
      }
      if (int e = launch_mlp_rows_backward(a.g, w.rows_x, w.tc_g, (int64_t)BN, w.gXg, &a.gg, st)) return e;
      if (t > 0) {
        PSVO_CUDA(cudaMemcpy2DAsync(w.rows_x, width, X_anc + (size_t)(t - 1) * N * Dx, pitch, width, B, cudaMemcpyDeviceToDevice, st));
        if (int e = launch_mlp_rows_forward(a.q1, w.rows_x, (int64_t)BN, w.tc_muq, st)) return e;
        a.ext_mu = w.tc_muq;
        {
          ProfScope ps_(PK_BWD, st);
          proposal_grad_kernel<<<prop_grid, 256, 0, st>>>(a);
          PSVO_LAUNCH_CHECK();
        }
        if (int e = launch_mlp_rows_backward(a.q1, w.rows_x, w.tc_g, (int64_t)BN, w.rows_gx, &a.gq1, st)) return e;
        {
          ProfScope ps_(PK_BWD, st);
          scatter_parent_kernel<<<light_grid, 256, 0, st>>>(w.rows_gx, anc + (size_t)(t - 1) * N, (long long)T * N, gXout, B, N, Dx);
          PSVO_LAUNCH_CHECK();
        }
      } else {
        a.ext_mu = nullptr; a.ext_g = nullptr;
        const int rc0 = DXP == 4 ? launch_proposal_t<4>(a, blkP, grid_of(blkP, smP), smP, st)
                                 : (DXP == 8 ? launch_proposal_t<8>(a, blkP, grid_of(blkP, smP), smP, st)
                                             : launch_proposal_t<16>(a, blkP, grid_of(blkP, smP), smP, st));
        if (rc0) return rc0;
      }
      float* tmp = gXin; gXin = gXout; gXout = tmp;
      continue;
    }
    int rc = 0;
    // small passes: the emission side depends on forward outputs only, so ONE launch covers all T steps (before the first step
    // of this loop); the per-step proposal kernels then read their slice of it
    const bool emis_all = w.gXg_all != nullptr;
    if (!emis_all || t == T - 1) {
      a.t_all = emis_all ? 1 : 0;
      a.gXg = emis_all ? w.gXg_all : w.gXg;
      const int gridE = emis_all ? (int)std::min<int64_t>((int64_t)grid_of(blkE, smE) * T, (int64_t)num_sms() * 4) : grid_of(blkE, smE);
      rc = -1;
#define PSVO_CASE_E(X, Y) if (DXP == X && DYP == Y) rc = launch_emission_t<X, Y>(a, blkE, gridE, smE, st);
      PSVO_CASE_E(4, 4) PSVO_CASE_E(4, 8) PSVO_CASE_E(4, 16) PSVO_CASE_E(4, 32)
      PSVO_CASE_E(8, 4) PSVO_CASE_E(8, 8) PSVO_CASE_E(8, 16) PSVO_CASE_E(8, 32)
      PSVO_CASE_E(16, 4) PSVO_CASE_E(16, 8) PSVO_CASE_E(16, 16) PSVO_CASE_E(16, 32)
#undef PSVO_CASE_E
      if (rc) return rc;
      a.t_all = 0;
    }
    if (emis_all) a.gXg = w.gXg_all + (size_t)t * BN * Dx;
    rc = DXP == 4 ? launch_proposal_t<4>(a, blkP, grid_of(blkP, smP), smP, st)
                  : (DXP == 8 ? launch_proposal_t<8>(a, blkP, grid_of(blkP, smP), smP, st)
                              : launch_proposal_t<16>(a, blkP, grid_of(blkP, smP), smP, st));
    if (rc) return rc;
    float* tmp = gXin; gXin = gXout; gXout = tmp;
  }

  // per-sequence networks: q2 over (b,t) rows, q0 (and f on pre_x0) over b rows
  if (two_q) {
    if (int e = launch_mlp_rows_backward(p->q2.net, pre_obs, w.gmu2_all, (int64_t)BT, g_pre_obs, &grads->q2.net, st)) return e;
  } else if (g_pre_obs != nullptr && pre_obs != nullptr) {
    PSVO_CUDA(cudaMemsetAsync(g_pre_obs, 0, BT * d->Dobs * sizeof(float), st));
  }
  if (int e = launch_mlp_rows_backward(p->q0.net, pre_x0, w.gmu0, B, g_pre_x0, &grads->q0.net, st)) return e;
  if (!f_is_d1) {
    const psvo_mlp_grad* gfd = boot ? &grads->q1.net : &grads->f.net;
    if (int e = launch_mlp_rows_backward(fd.net, pre_x0, w.gmuf0, B, g_pre_x0 ? w.gx0tmp : nullptr, gfd, st)) return e;
    if (g_pre_x0 != nullptr) {
      ProfScope ps_(PK_BWD, st);
      add_inplace_kernel<<<ceil_div((int64_t)B * d->Dx0, 256), 256, 0, st>>>(g_pre_x0, w.gx0tmp, (long long)B * d->Dx0);
      PSVO_LAUNCH_CHECK();
    }
  }

  SigmaGradArgs sa;
  sa.stats = w.stats; sa.sig = w.sig; sa.Dx = Dx; sa.Dy = d->Dy; sa.flags = d->flags;
  const SigmaSrc src = sigma_src(d, p, false);
  for (int k = 0; k < SIG_SLOTS; ++k) { sa.con[k] = src.con[k]; sa.smin[k] = src.smin[k]; sa.gcon[k] = nullptr; }
  sa.gcon[SIG_Q0] = grads->q0.sigma_con; sa.gcon[SIG_Q1] = grads->q1.sigma_con; sa.gcon[SIG_Q2] = grads->q2.sigma_con;
  sa.gcon[SIG_F] = grads->f.sigma_con; sa.gcon[SIG_G] = grads->g.sigma_con;
  {
    ProfScope ps_(PK_BWD, st);
    sigma_grad_kernel<<<1, 32, 0, st>>>(sa);
    PSVO_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
