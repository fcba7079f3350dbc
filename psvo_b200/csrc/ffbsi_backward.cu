// ffbsi_backward.cu -- reverse mode of the FFBSi backward simulation (PSVO) and of its "with resampling" variant.
//
// Reference: what optimizer.minimize(-log_ZSMC) (src/trainer.py:115-118) derives through
// PSVO.backward_simulation_w_proposal + compute_log_ZSMC (src/SMC/PSVO.py:52-203) and the PSVOwR variants
// (src/SMC/PSVOwR.py:52-198).  Everything except the index draws is differentiable there: the reparameterised
// candidates x~ = mean + s*eps, the gather of the picked candidate (and PSVOwR's second gather), the three MLP
// log densities, the PAIRWISE predictive term  P[m,i] = LSE_j( log f(x~[m,i] | Xs[t-1][j]) + log Wbar_{t-1}[j] )
// and the normalisation over the M candidates.  Nothing flows into the 1-of-M pick / resampling logits
// (tf.multinomial has no gradient).  Gradients w.r.t. the forward filter's Xs and log_Ws come back as g_Xs /
// g_log_Ws and are fed to psvo_smc_backward.
//
// With pi_m = softmax_m(omega), delta = [m is the picked candidate] and c = d loss / d (trajectory term):
//   PSVO   S_t,i = LSE_m omega - P* - log M  (+ P* at t = 0), c_i = g_logZ_b * softmax_i(sum_t S_t,i):
//            coefficient of P_m: (pi_m - delta) c   (pi_m c at t = 0) ; of f_m, g_m: pi_m c ; of q_m: -pi_m c
//   PSVOwR bw_log_W_t,i = LSE_m omega - f* - log M, c_t,i = g_logZ_b * softmax_i(bw_log_W_t):
//            coefficient of P_m, g_m: pi_m c ; of f_m: (pi_m - delta) c ; of q_m: -pi_m c
// q_m depends on the parameters only through -sum log s (x~ - mean = s*eps exactly).
//
// The simulation ran t = T-1 .. 0, so its reverse runs t = 0 .. T-1.  Per step (candidates are recomputed from
// eps_bw / Philox and the stored picks, never stored by the forward pass):
//   ffbsi_bwd_cand : thread per candidate (m,i): recompute x~, f, g, P (pairwise pass 1), pi; back-propagate g and f
//                    through their MLPs, the pairwise term w.r.t. x~ (pass 2), the initial density at t = 0; add the
//                    gradient arriving at the picked candidate from step t-1; reduce over m to the gradient of the
//                    proposal mean; park (x~/s_f, kappa, P) for the j-side kernel.
//   ffbsi_bwd_pair : (t >= 1) thread per filter particle j of Xs[t-1]: loop over all M*N candidates of its
//                    sequence, accumulate d/d mu_f[j] and d/d log Wbar[j]; MLP f reverse -> g_Xs[t-1], dW_f.
//   ffbsi_bwd_traj : (t <= T-2) thread per trajectory: MLP q1_inv reverse -> gradient w.r.t. x_{t+1}^{bw}, handed
//                    to the picked candidate of step t+1 (through PSVOwR's second gather when present).
#include <map>
#include <mutex>
#include "mlp_bwd_tile.cuh"
#include "smc_kernels.h"

namespace psvo {

int check_dims(const psvo_dims* d);
int check_mlp(const char* name, const psvo_mlp& net, int d_in, int d_out);
SigmaSrc sigma_src(const psvo_dims* d, const psvo_params* p, bool ffbsi);

enum { FS_SP_MID = 0, FS_I1, FS_I2, FS_SP_LAST, FS_SF, FS_SG, FS_SINIT, FS_SLOTS };
constexpr int kBwdPairChunk = 128;

struct FbArgs {
  psvo_mlp f, g, q1i;
  psvo_mlp_grad gf, gg, gq1i;
  const float* sig;
  const float* obs; const float* Xs; const float* log_Ws; const float* lse;
  const float* mu_f; const float* mu1i; const float* m2_all; const float* mu_init_all; const float* mu0i;
  const float* X_next;       // (B,T,N,Dx): bw_Xs (PSVO) or bw_X_anc (PSVOwR); slab t+1 is read
  const float* eps_bw; const int* pick; const int* anc2;
  const float* coef;         // PSVO (B,N) ; PSVOwR (B,T,N)
  const float* gin;          // (B,N,Dx) gradient arriving at the picked candidate of trajectory i
  float* gin_next;           // (B,N,Dx) produced by ffbsi_bwd_traj for step t+1
  float* candX; float* candK; float* candP;       // (B,N,M,DXP) | (B,N,M) | (B,N,M)
  int pair_all = 0;          // ffbsi_bwd_pair_kernel only: 1 = the j side of ALL steps t = 1..T-1 in one launch (batch 1: candX / candK /
                             // candP hold every step's candidates, mu_f points at the (T,N,Dx) means of all steps)
  float* gmu1i; float* gxn;  // (B,N,Dx)
  float* gm2_all; float* gmu_init_all; float* gmu0;
  float* g_Xs; float* g_log_Ws;
  float* stats;
  unsigned long long seed;
  int N, M, B, T, Dx, Dy, t, seq_offset, IL;
  unsigned flags;
};

__device__ __forceinline__ void warp_acc(float v, float* dst) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0 && v != 0.0f) atomicAdd(dst, v);
}

// ------------------------------------------------------------------------------------------------------
// candidate-side kernel
// ------------------------------------------------------------------------------------------------------
template <int DXP, int DYP>
__global__ void __launch_bounds__(256) ffbsi_bwd_cand_kernel(FbArgs a) {
  griddep_entry();
  extern __shared__ __align__(16) float smem[];
  const int Dx = a.Dx, Dy = a.Dy, N = a.N, M = a.M, T = a.T, t = a.t;
  const bool boot = a.flags & PSVO_F_BOOTSTRAP, two_q = a.flags & PSVO_F_TWO_Q, poisson = a.flags & PSVO_F_POISSON;
  const bool wr = a.flags & PSVO_F_WITH_RESAMPLING;
  const bool last = t == T - 1, first = t == 0;
  const int P = blockDim.x, tid = threadIdx.x;

  MlpS<DXP, DXP> fs;
  MlpS<DXP, DYP> gs;
  MlpGradS<DXP, DXP> gfS;
  MlpGradS<DXP, DYP> ggS;
  float* p = smem;
  float* sp = p; p += DXP;  float* inv1 = p; p += DXP;  float* inv2 = p; p += DXP;
  float* sf = p; p += DXP;  float* sinit = p; p += DXP; float* sg = p; p += 32;
  float* lds = p; p += 4;
  float* sred = p; p += 2 * 32;               // block partials: gm2 (mid) or gmu_init (last) | gmu0 (first)
  float* sstat = p; p += FS_SLOTS * 32;
  if (!last) p = mlp_stage<DXP, DXP>(a.f, p, fs);
  p = mlp_stage<DXP, DYP>(a.g, p, gs);
  float* gacc0 = p;
  if (!last) p = mlp_grad_carve<DXP, DXP>(a.f, p, gfS);
  p = mlp_grad_carve<DXP, DYP>(a.g, p, ggS);
  for (float* q = gacc0 + tid; q < p; q += P) *q = 0.0f;
  float* sMu = p; p += kBwdPairChunk * DXP;
  float* sWb = p; p += kBwdPairChunk;
  float* sOm = p; p += P;
  float* sGX = p; p += (size_t)P * 2 * DXP;
  float* stage = p;

  if (tid < DXP) {
    const int d = tid;
    const float s1 = a.sig[(last ? SIG_BQI : SIG_Q1I) * 32 + d], s2 = a.sig[SIG_BQ2 * 32 + d];
    inv1[d] = 1.0f / s1; inv2[d] = 1.0f / s2;
    sp[d] = last ? s1 : 1.0f / (1.0f / s1 + 1.0f / s2);
    sf[d] = a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
    sinit[d] = a.sig[((boot && two_q) ? SIG_Q0 : (boot ? SIG_Q1 : SIG_F)) * 32 + d];
  }
  for (int i = tid; i < 32; i += P) sg[i] = a.sig[SIG_G * 32 + i];
  for (int i = tid; i < FS_SLOTS * 32; i += P) sstat[i] = 0.0f;
  for (int i = tid; i < 2 * 32; i += P) sred[i] = 0.0f;
  __syncthreads();
  if (tid == 0) {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int d = 0; d < Dx; ++d) { l0 += logf(sp[d]); l1 += logf(sf[d]); l3 += logf(sinit[d]); }
    if (!poisson) for (int d = 0; d < Dy; ++d) l2 += logf(sg[d]);
    lds[0] = l0; lds[1] = l1; lds[2] = l2; lds[3] = l3;
  }
  __syncthreads();
  const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;

  const int IL = a.IL;
  const int il = tid / M, m = tid - il * M;
  const int tiles_per_seq = (N + IL - 1) / IL;
  const int work = a.B * tiles_per_seq;

  for (int wi = blockIdx.x; wi < work; wi += gridDim.x) {
    const int b = wi / tiles_per_seq;
    const int i = (wi - b * tiles_per_seq) * IL + il;
    const bool live = il < IL && i < N;
    const size_t bt = (size_t)b * T + t;
    const int ii = live ? i : 0;
    const size_t bi = (size_t)b * N + ii;

    // ---- recompute the candidate  x~ = mean + sp * eps
    float X[1][DXP], e[DXP], xn[DXP], Am[DXP];
#pragma unroll
    for (int c = 0; c < DXP / 4; ++c) {
      float ev[4] = {0.f, 0.f, 0.f, 0.f};
      if (4 * c < Dx) {
        if (a.eps_bw != nullptr) {
          const float* src = a.eps_bw + (((bt * M + (live ? m : 0)) * N + ii) * (size_t)Dx);
#pragma unroll
          for (int k = 0; k < 4; ++k) if (4 * c + k < Dx) ev[k] = src[4 * c + k];
        } else {
          const uint64_t gidx = ((((uint64_t)(a.seq_offset + b) * T + t) * M + (live ? m : 0)) * N) + ii;
          normal4(philox4x32_10(a.seed, gidx, c, 2u), ev);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = 4 * c + k;
        float mean = 0.f, xnext = 0.f, am = 0.f;
        if (d < Dx) {
          if (last) mean = a.mu_init_all[bt * Dx + d];
          else {
            am = inv1[d] * a.mu1i[bi * Dx + d] + inv2[d] * a.m2_all[bt * Dx + d];
            mean = sp[d] * am;
            xnext = a.X_next[((bt + 1) * N + ii) * Dx + d];
          }
        }
        X[0][d] = d < Dx ? mean + sp[d] * ev[k] : 0.0f;
        e[d] = d < Dx ? ev[k] : 0.0f;
        xn[d] = xnext; Am[d] = am;
      }
    }
    float qs = 0.f;
#pragma unroll
    for (int d = 0; d < DXP; ++d) qs = fmaf(e[d], e[d], qs);          // (x~ - mean) / sp = eps
    const float q_lp = -0.5f * qs - lds[0] - cx;

    // ---- f(x_{t+1} | x~) and g(y_t | x~)
    float zf[DXP];
    float f_lp = 0.0f;
    if (!last) {
      float muf[1][DXP];
      mlp_eval<DXP, DXP, 1>(fs, X, muf, nullptr, 0);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DXP; ++d) { zf[d] = d < Dx ? (xn[d] - muf[0][d]) / sf[d] : 0.0f; s = fmaf(zf[d], zf[d], s); }
      f_lp = -0.5f * s - lds[1] - cx;
    } else {
#pragma unroll
      for (int d = 0; d < DXP; ++d) zf[d] = 0.0f;
    }
    float G[DYP];           // d g_lp / d mug
    float zg2[DYP];         // z^2 (Gaussian emission), for the sigma_g statistic
    float g_lp;
    {
      float mug[1][DYP];
      mlp_eval<DXP, DYP, 1>(gs, X, mug, nullptr, 0);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DYP; ++d) {
        G[d] = 0.0f; zg2[d] = 0.0f;
        if (d < Dy) {
          const float y = a.obs[bt * Dy + d];
          if (poisson) {
            const float z = y - (softplus(mug[0][d]) + 1e-6f);
            G[d] = z * sigmoidf(mug[0][d]);
            s = fmaf(z, z, s);
          } else {
            const float z = (y - mug[0][d]) / sg[d];
            G[d] = z / sg[d]; zg2[d] = z * z;
            s = fmaf(z, z, s);
          }
        }
      }
      g_lp = -0.5f * s - lds[2] - cy;
    }

    // ---- predictive term (pass 1: log-sum-exp over j)
    float xs[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) xs[d] = d < Dx ? X[0][d] / sf[d] : 0.0f;
    float Praw = 0.0f, Pv;
    float zi[DXP];
    if (!first) {
      float mx = -INFINITY, sum = 0.0f;
      const float lse_prev = a.lse[bt - 1];
      for (int j0 = 0; j0 < N; j0 += kBwdPairChunk) {
        __syncthreads();
        for (int q = tid; q < kBwdPairChunk * DXP; q += P) {
          const int jj = q / DXP, d = q - jj * DXP;
          sMu[q] = (j0 + jj < N && d < Dx) ? a.mu_f[((size_t)b * N + j0 + jj) * Dx + d] / sf[d] : 0.0f;
        }
        for (int q = tid; q < kBwdPairChunk; q += P) sWb[q] = j0 + q < N ? a.log_Ws[(bt - 1) * N + j0 + q] - lse_prev : -INFINITY;
        __syncthreads();
        const int nj = min(kBwdPairChunk, N - j0);
        for (int jj = 0; jj < nj; ++jj) {
          float muv[DXP];
          lds_vec<DXP>(sMu + jj * DXP, muv);
          float d2 = 0.f;
#pragma unroll
          for (int d = 0; d < DXP; ++d) { const float z = xs[d] - muv[d]; d2 = fmaf(z, z, d2); }
          const float v = fmaf(-0.5f, d2, sWb[jj]);
          if (v > mx) { sum = sum * __expf(mx - v) + 1.0f; mx = v; }
          else sum += __expf(v - mx);
        }
      }
      Praw = mx + logf(sum);
      Pv = Praw - lds[1] - cx;
    } else {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DXP; ++d) { zi[d] = d < Dx ? (X[0][d] - a.mu0i[(size_t)b * Dx + d]) / sinit[d] : 0.0f; s = fmaf(zi[d], zi[d], s); }
      Pv = -0.5f * s - lds[3] - cx;
    }
    const float omega = last ? (Pv + g_lp) - q_lp : ((Pv + f_lp) + g_lp) - q_lp;

    // ---- pi_m = softmax over the M candidates of this trajectory; coefficients
    __syncthreads();
    sOm[tid] = omega;
    __syncthreads();
    float pi = 0.0f, c = 0.0f, dl = 0.0f;
    if (live) {
      const float* om = sOm + il * M;
      float mxo = -INFINITY;
      for (int k = 0; k < M; ++k) mxo = fmaxf(mxo, om[k]);
      float se = 0.f;
      for (int k = 0; k < M; ++k) se += expf(om[k] - mxo);
      pi = expf(omega - (mxo + logf(se)));
      c = wr ? a.coef[bt * N + i] : a.coef[(size_t)b * N + i];
      dl = (m == a.pick[bt * N + i]) ? 1.0f : 0.0f;
    }
    const float cP = (wr || first) ? pi * c : (pi - dl) * c;
    const float cf = wr ? (pi - dl) * c : pi * c;
    const float cg = pi * c;

    float gX[DXP], gx[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) gX[d] = 0.0f;

    // ---- emission
#pragma unroll
    for (int d = 0; d < DYP; ++d) {
      if (!poisson && d < Dy) warp_acc(cg * (zg2[d] - 1.0f) / sg[d], &sstat[FS_SG * 32 + d]);
      G[d] *= cg;
    }
    mlp_bwd_tile<DXP, DYP>(gs, ggS, X[0], G, gx, stage);
#pragma unroll
    for (int d = 0; d < DXP; ++d) gX[d] += gx[d];

    // ---- transition to the next state
    float gxn[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) gxn[d] = 0.0f;
    if (!last) {
      float gmuf[DXP];
#pragma unroll
      for (int d = 0; d < DXP; ++d) {
        gmuf[d] = d < Dx ? cf * zf[d] / sf[d] : 0.0f;
        gxn[d] = -gmuf[d];
        if (d < Dx) warp_acc(cf * (zf[d] * zf[d] - 1.0f) / sf[d], &sstat[FS_SF * 32 + d]);
      }
      mlp_bwd_tile<DXP, DXP>(fs, gfS, X[0], gmuf, gx, stage);
#pragma unroll
      for (int d = 0; d < DXP; ++d) gX[d] += gx[d];
    }

    // ---- predictive term (pass 2: softmax-weighted differences) / initial density
    if (!first) {
      float a0 = 0.0f, a1[DXP], a2[DXP];
#pragma unroll
      for (int d = 0; d < DXP; ++d) a1[d] = a2[d] = 0.0f;
      const float lse_prev = a.lse[bt - 1];
      for (int j0 = 0; j0 < N; j0 += kBwdPairChunk) {
        __syncthreads();
        for (int q = tid; q < kBwdPairChunk * DXP; q += P) {
          const int jj = q / DXP, d = q - jj * DXP;
          sMu[q] = (j0 + jj < N && d < Dx) ? a.mu_f[((size_t)b * N + j0 + jj) * Dx + d] / sf[d] : 0.0f;
        }
        for (int q = tid; q < kBwdPairChunk; q += P) sWb[q] = j0 + q < N ? a.log_Ws[(bt - 1) * N + j0 + q] - lse_prev : -INFINITY;
        __syncthreads();
        const int nj = min(kBwdPairChunk, N - j0);
        for (int jj = 0; jj < nj; ++jj) {
          float muv[DXP];
          lds_vec<DXP>(sMu + jj * DXP, muv);
          float d2 = 0.f;
#pragma unroll
          for (int d = 0; d < DXP; ++d) { muv[d] = xs[d] - muv[d]; d2 = fmaf(muv[d], muv[d], d2); }
          const float rho = __expf(fmaf(-0.5f, d2, sWb[jj]) - Praw);
          a0 += rho;
#pragma unroll
          for (int d = 0; d < DXP; ++d) { a1[d] = fmaf(rho, muv[d], a1[d]); a2[d] = fmaf(rho * muv[d], muv[d], a2[d]); }
        }
      }
#pragma unroll
      for (int d = 0; d < DXP; ++d) {
        if (d < Dx) {
          gX[d] -= cP * a1[d] / sf[d];
          warp_acc(cP * (a2[d] - a0) / sf[d], &sstat[FS_SF * 32 + d]);
        }
      }
      if (live) {
        const size_t ci = ((size_t)b * N + i) * M + m;
        a.candK[ci] = cP; a.candP[ci] = Praw;
#pragma unroll
        for (int d = 0; d < DXP; ++d) a.candX[ci * DXP + d] = xs[d];
      }
    } else {
#pragma unroll
      for (int d = 0; d < DXP; ++d) {
        if (d < Dx) {
          const float gm = cP * zi[d] / sinit[d];
          gX[d] -= gm;
          warp_acc(gm, &sred[32 + d]);
          warp_acc(cP * (zi[d] * zi[d] - 1.0f) / sinit[d], &sstat[FS_SINIT * 32 + d]);
        }
      }
    }

    // ---- gradient arriving at the picked candidate from step t-1
    if (live && dl != 0.0f && a.gin != nullptr) {
#pragma unroll
      for (int d = 0; d < DXP; ++d) if (d < Dx) gX[d] += a.gin[bi * Dx + d];
    }

    // ---- reparameterisation  x~ = sp * (inv1 mu1 + inv2 m2) + sp * eps   (last step: mu_init + s eps)
    const float piq = pi * c;         // -q_m carries +pi c / sp per dimension
#pragma unroll
    for (int d = 0; d < DXP; ++d) {
      if (d < Dx) {
        const float st = gX[d] * (e[d] + Am[d]) + piq / sp[d];
        warp_acc(st, &sstat[(last ? FS_SP_LAST : FS_SP_MID) * 32 + d]);
        if (!last) {
          const float m1 = live ? a.mu1i[bi * Dx + d] : 0.0f, m2 = a.m2_all[bt * Dx + d];
          warp_acc(gX[d] * sp[d] * m1, &sstat[FS_I1 * 32 + d]);
          warp_acc(gX[d] * sp[d] * m2, &sstat[FS_I2 * 32 + d]);
          warp_acc(gX[d] * sp[d] * inv2[d], &sred[d]);
        } else {
          warp_acc(gX[d], &sred[d]);
        }
      }
      sGX[(size_t)tid * 2 * DXP + d] = gX[d];
      sGX[(size_t)tid * 2 * DXP + DXP + d] = gxn[d];
    }
    __syncthreads();
    if (live && m == 0 && !last) {
      const float* src = sGX + (size_t)il * M * 2 * DXP;
      for (int d = 0; d < Dx; ++d) {
        float gm = 0.0f, gn = 0.0f;
        for (int k = 0; k < M; ++k) { gm += src[(size_t)k * 2 * DXP + d]; gn += src[(size_t)k * 2 * DXP + DXP + d]; }
        a.gmu1i[bi * Dx + d] = gm * sp[d] * inv1[d];
        a.gxn[bi * Dx + d] = gn;
      }
    }
    if (tid < Dx) {
      const int d = tid;
      if (last) atomicAdd(a.gmu_init_all + bt * Dx + d, sred[d]);
      else atomicAdd(a.gm2_all + bt * Dx + d, sred[d]);
      sred[d] = 0.0f;
      if (first) { atomicAdd(a.gmu0 + (size_t)b * Dx + d, sred[32 + d]); sred[32 + d] = 0.0f; }
    }
    __syncthreads();
  }

  __syncthreads();
  for (int i = tid; i < FS_SLOTS * 32; i += P) { const float v = sstat[i]; if (v != 0.0f) atomicAdd(a.stats + i, v); }
  if (!last) mlp_grad_flush<DXP, DXP>(a.f, gfS, a.gf);
  mlp_grad_flush<DXP, DYP>(a.g, ggS, a.gg);
}

// ------------------------------------------------------------------------------------------------------
// j-side kernel of the pairwise term  (t >= 1)
// ------------------------------------------------------------------------------------------------------
template <int DXP>
__global__ void __launch_bounds__(256) ffbsi_bwd_pair_kernel(FbArgs a) {
  griddep_entry();
  extern __shared__ __align__(16) float smem[];
  const int Dx = a.Dx, N = a.N, M = a.M, T = a.T;
  const bool boot = a.flags & PSVO_F_BOOTSTRAP;
  const int P = blockDim.x, tid = threadIdx.x;
  MlpS<DXP, DXP> fs;
  MlpGradS<DXP, DXP> gfS;
  float* p = smem;
  float* sf = p; p += DXP;
  p = mlp_stage<DXP, DXP>(a.f, p, fs);
  float* gacc0 = p;
  p = mlp_grad_carve<DXP, DXP>(a.f, p, gfS);
  for (float* q = gacc0 + tid; q < p; q += P) *q = 0.0f;
  float* sCX = p; p += kBwdPairChunk * DXP;
  float* sCK = p; p += kBwdPairChunk;
  float* sCP = p; p += kBwdPairChunk;
  float* stage = p;
  if (tid < DXP) sf[tid] = a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + tid];
  __syncthreads();

  // few particles (2 N <= P): the block splits the candidate loop of every particle j over S = P / NP2 slices (thread = (slice,
  // j)), partial sums meet in shared memory in a fixed order, and the slice-0 threads carry on as the particles' owners
  int NP2 = P, S = 1;
  if (2 * N <= P) { NP2 = 1; while (NP2 < N) NP2 *= 2; S = P / NP2; }
  const int slice = tid / NP2;
  const int tiles_per_seq = S > 1 ? 1 : (N + P - 1) / P;
  const int work_t = a.B * tiles_per_seq;
  const int work = a.pair_all ? work_t * (T - 1) : work_t;
  const long long ncand = (long long)N * M;
  for (int wj = blockIdx.x; wj < work; wj += gridDim.x) {
    const int t = a.pair_all ? 1 + wj / work_t : a.t;
    const int wi = a.pair_all ? wj - (t - 1) * work_t : wj;
    const int b = wi / tiles_per_seq;
    const int j = S > 1 ? tid % NP2 : (wi - b * tiles_per_seq) * P + tid;
    const bool live = j < N;
    const size_t btm = (size_t)b * T + (t - 1);
    const size_t step_off = a.pair_all ? (size_t)t * a.B * ncand : 0;               // this step's slice of the candidate buffers
    const float* mu_f_t = a.pair_all ? a.mu_f + (size_t)(t - 1) * N * Dx : a.mu_f;  // (batch 1)
    float x[DXP], mus[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) {
      x[d] = (live && d < Dx) ? a.Xs[(btm * N + j) * Dx + d] : 0.0f;
      mus[d] = (live && d < Dx) ? mu_f_t[((size_t)b * N + j) * Dx + d] / sf[d] : 0.0f;
    }
    const float wb = live ? a.log_Ws[btm * N + j] - a.lse[btm] : -INFINITY;
    float Gj = 0.0f, Ksum = 0.0f, gm[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) gm[d] = 0.0f;
    const float* cX = a.candX + (step_off + (size_t)b * ncand) * DXP;
    const float* cK = a.candK + step_off + (size_t)b * ncand;
    const float* cP = a.candP + step_off + (size_t)b * ncand;
    for (long long c0 = 0; c0 < ncand; c0 += kBwdPairChunk) {
      __syncthreads();
      const int nc = (int)min((long long)kBwdPairChunk, ncand - c0);
      for (int q = tid; q < nc * DXP; q += P) sCX[q] = cX[c0 * DXP + q];
      for (int q = tid; q < nc; q += P) { sCK[q] = cK[c0 + q]; sCP[q] = cP[c0 + q]; }
      __syncthreads();
      for (int cc = slice; cc < nc; cc += S) {
        float dv[DXP];
        lds_vec<DXP>(sCX + cc * DXP, dv);
        float d2 = 0.f;
#pragma unroll
        for (int d = 0; d < DXP; ++d) { dv[d] -= mus[d]; d2 = fmaf(dv[d], dv[d], d2); }
        const float w = sCK[cc] * __expf(fmaf(-0.5f, d2, wb) - sCP[cc]);
        Gj += w; Ksum += sCK[cc];
#pragma unroll
        for (int d = 0; d < DXP; ++d) gm[d] = fmaf(w, dv[d], gm[d]);
      }
    }
    if (S > 1) {
      // partial sums of the slices -> the owner (slice 0), summed in slice order; the staging area is free until mlp_bwd_tile
      __syncthreads();
      float* part = stage + (size_t)tid * (2 + DXP);
      part[0] = Gj; part[1] = Ksum;
#pragma unroll
      for (int d = 0; d < DXP; ++d) part[2 + d] = gm[d];
      __syncthreads();
      if (slice == 0) {
        for (int sl = 1; sl < S; ++sl) {
          const float* q = stage + (size_t)(sl * NP2 + tid) * (2 + DXP);
          Gj += q[0]; Ksum += q[1];
#pragma unroll
          for (int d = 0; d < DXP; ++d) gm[d] += q[2 + d];
        }
      }
    }
    const bool owner = live && slice == 0;
    float gmu[DXP], gx[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) gmu[d] = (owner && d < Dx) ? gm[d] / sf[d] : 0.0f;
    // log Wbar[j] = log_W[j] - LSE_j log_W: the normaliser sends -softmax_j * sum_j' G_j' back, and sum_j' G_j' is
    // the sum of kappa over the candidates (sum_j rho = 1).  It vanishes for PSVO (sum_m (pi_m - delta) = 0), not for
    // PSVOwR (kappa = pi_m c).
    if (owner) a.g_log_Ws[btm * N + j] = Gj - expf(wb) * Ksum;
    mlp_bwd_tile<DXP, DXP>(fs, gfS, x, gmu, gx, stage);
    if (owner) {
#pragma unroll
      for (int d = 0; d < DXP; ++d) if (d < Dx) a.g_Xs[(btm * N + j) * Dx + d] = gx[d];
    }
  }
  __syncthreads();
  mlp_grad_flush<DXP, DXP>(a.f, gfS, a.gf);
}

// ------------------------------------------------------------------------------------------------------
// trajectory kernel: MLP q1_inv reverse  (t <= T-2)
// ------------------------------------------------------------------------------------------------------
template <int DXP>
__global__ void __launch_bounds__(256) ffbsi_bwd_traj_kernel(FbArgs a) {
  griddep_entry();
  extern __shared__ __align__(16) float smem[];
  const int Dx = a.Dx, N = a.N, T = a.T, t = a.t;
  const bool wr = a.flags & PSVO_F_WITH_RESAMPLING;
  const int P = blockDim.x, tid = threadIdx.x;
  MlpS<DXP, DXP> qs;
  MlpGradS<DXP, DXP> gqS;
  float* p = smem;
  p = mlp_stage<DXP, DXP>(a.q1i, p, qs);
  float* gacc0 = p;
  p = mlp_grad_carve<DXP, DXP>(a.q1i, p, gqS);
  for (float* q = gacc0 + tid; q < p; q += P) *q = 0.0f;
  float* stage = p;
  __syncthreads();
  const int tiles_per_seq = (N + P - 1) / P;
  const int work = a.B * tiles_per_seq;
  for (int wi = blockIdx.x; wi < work; wi += gridDim.x) {
    const int b = wi / tiles_per_seq;
    const int i = (wi - b * tiles_per_seq) * P + tid;
    const bool live = i < N;
    const size_t btn = (size_t)b * T + (t + 1);
    const size_t bi = (size_t)b * N + (live ? i : 0);
    float x[DXP], go[DXP], gx[DXP];
#pragma unroll
    for (int d = 0; d < DXP; ++d) {
      x[d] = (live && d < Dx) ? a.X_next[(btn * N + i) * Dx + d] : 0.0f;
      go[d] = (live && d < Dx) ? a.gmu1i[bi * Dx + d] : 0.0f;
    }
    mlp_bwd_tile<DXP, DXP>(qs, gqS, x, go, gx, stage);
    if (live) {
      if (wr) {
        const int parent = a.anc2[btn * N + i];
        float* dst = a.gin_next + ((size_t)b * N + parent) * Dx;
#pragma unroll
        for (int d = 0; d < DXP; ++d) if (d < Dx) atomicAdd(dst + d, gx[d] + a.gxn[bi * Dx + d]);
      } else {
#pragma unroll
        for (int d = 0; d < DXP; ++d) if (d < Dx) a.gin_next[bi * Dx + d] = gx[d] + a.gxn[bi * Dx + d];
      }
    }
  }
  __syncthreads();
  mlp_grad_flush<DXP, DXP>(a.q1i, gqS, a.gq1i);
}

// ------------------------------------------------------------------------------------------------------
// coefficients c = d loss / d (trajectory term)
//   PSVO   (grid B)  : coef[b][i]    = g_logZ_b[b] * softmax_i( sum_t f + g - Omega )
//   PSVOwR (grid B*T): coef[b][t][i] = g_logZ_b[b] * softmax_i( bw_log_W[b][t][:] )
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffbsi_coef_kernel(const float* __restrict__ f_lp, const float* __restrict__ g_lp,
                                                         const float* __restrict__ oa, const float* __restrict__ g_logZ_b,
                                                         int N, int T, int wr, float* __restrict__ coef) {
  __shared__ float scr[33];
  const int b = wr ? blockIdx.x / T : blockIdx.x;
  float* row = coef + (size_t)blockIdx.x * N;
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float v;
    if (wr) v = oa[(size_t)blockIdx.x * N + n];
    else {
      float joint = 0.f, prop = 0.f;
      for (int t = 0; t < T; ++t) {
        const size_t k = ((size_t)b * T + t) * N + n;
        joint += f_lp[k] + g_lp[k];
        prop += oa[k];
      }
      v = joint - prop;
    }
    row[n] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, scr);
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(row[n] - mx);
  s = block_sum(s, scr);
  const float g = g_logZ_b[b] / s;
  for (int n = threadIdx.x; n < N; n += blockDim.x) row[n] = expf(row[n] - mx) * g;
}

// statistics -> d/d sigma_con
struct FbSigmaArgs {
  const float* stats; const float* sig;
  const float* con[SIG_SLOTS]; float smin[SIG_SLOTS]; float* gcon[SIG_SLOTS];
  int Dx, Dy; unsigned flags;
};
__global__ void ffbsi_sigma_grad_kernel(FbSigmaArgs a) {
  const int d = threadIdx.x;
  const bool two_q = a.flags & PSVO_F_TWO_Q, boot = a.flags & PSVO_F_BOOTSTRAP, poisson = a.flags & PSVO_F_POISSON;
  float gs[SIG_SLOTS];
  for (int k = 0; k < SIG_SLOTS; ++k) gs[k] = 0.0f;
  if (d < a.Dx) {
    const float s1 = a.sig[SIG_Q1I * 32 + d], s2 = a.sig[SIG_BQ2 * 32 + d];
    const float i1 = 1.0f / s1, i2 = 1.0f / s2, sp = 1.0f / (i1 + i2);
    const float acc = a.stats[FS_SP_MID * 32 + d];
    gs[SIG_Q1I] += -(a.stats[FS_I1 * 32 + d] - sp * sp * acc) * i1 * i1;
    gs[SIG_BQ2] += -(a.stats[FS_I2 * 32 + d] - sp * sp * acc) * i2 * i2;
    gs[SIG_BQI] += a.stats[FS_SP_LAST * 32 + d];
    gs[boot ? SIG_Q1 : SIG_F] += a.stats[FS_SF * 32 + d];
    gs[(boot && two_q) ? SIG_Q0 : (boot ? SIG_Q1 : SIG_F)] += a.stats[FS_SINIT * 32 + d];
  }
  if (!poisson && d < a.Dy) gs[SIG_G] += a.stats[FS_SG * 32 + d];
  for (int k = 0; k < SIG_SLOTS; ++k) {
    const int dim = k == SIG_G ? a.Dy : a.Dx;
    if (a.gcon[k] == nullptr || a.con[k] == nullptr || d >= dim) continue;
    const float sc = a.con[k][d];
    const float sval = softplus(sc);
    const bool pass = (sc == sc) && (sval > a.smin[k]);
    if (pass) atomicAdd(a.gcon[k] + d, gs[k] * sigmoidf(sc));
  }
}

__global__ void fb_add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}

static size_t staged_fl(const psvo_mlp& net, int din_pad, int dout_pad) {
  if (net.n_hidden == 0) return (size_t)din_pad * dout_pad + dout_pad;
  return (size_t)net.width[0] * din_pad + round4(net.width[0]) + (size_t)net.width[0] * dout_pad + dout_pad;
}

template <typename K>
static int launch_fb(K kern, const FbArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
  PSVO_DYN_SMEM(kern, smem);      // once per (device, kernel): a driver call per launch would dominate small steps
  ProfScope ps_(PK_FFBSI, st);
  PSVO_CUDA(launch_pdl(kern, dim3(grid), dim3(block), smem, st, pdl_enabled(), a));
  PSVO_LAUNCH_CHECK();
  return 0;
}

struct FbWs {
  int cand_steps;                  // slices of candX / candK / candP (T when the pair kernel runs once per pass, else 1)
  float *mu_f_all, *mu1i_all;      // batch 1: the means of ALL time steps, two launches per pass instead of two per step (or NULL)
  float *sig, *mu_f, *mu1i, *m2_all, *mu_init_all, *mu0i, *coef, *candX, *candK, *candP, *gmu1i, *gxn, *ginA, *ginB,
        *gm2_all, *gmu_init_all, *gmu0, *stats, *gfeat_tmp, *gx0_tmp;
  size_t bytes;
};
static FbWs carve_fb(const psvo_dims* d, char* base) {
  FbWs w; size_t off = 0;
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(base + off); off += align_up(floats * sizeof(float), 256); return r; };
  const size_t BT = (size_t)d->B * d->T, BN = (size_t)d->B * d->N, BNM = BN * d->M;
  const int DXP = pad_dx(d->Dx);
  const bool wr = d->flags & PSVO_F_WITH_RESAMPLING;
  w.sig = take(SIG_SLOTS * 32);
  w.mu_f = take(BN * d->Dx); w.mu1i = take(BN * d->Dx);
  const bool all_steps = d->B == 1 && (size_t)d->T * d->N * d->Dx <= (size_t)1 << 26;
  w.mu_f_all = all_steps ? take((size_t)d->T * d->N * d->Dx) : nullptr;
  w.mu1i_all = all_steps ? take((size_t)d->T * d->N * d->Dx) : nullptr;
  w.m2_all = take(BT * d->Dx); w.mu_init_all = take(BT * d->Dx); w.mu0i = take((size_t)d->B * d->Dx);
  w.coef = take(wr ? BT * d->N : BN);
  // batch 1 with small candidate sets: one slice per time step, so that the j side of the pairwise term (it only feeds outputs
  // and the f-network gradients) runs for ALL steps in one launch after the loop
  w.cand_steps = (all_steps && (size_t)d->T * BNM * (DXP + 2) <= (size_t)1 << 23) ? d->T : 1;
  w.candX = take(w.cand_steps * BNM * DXP); w.candK = take(w.cand_steps * BNM); w.candP = take(w.cand_steps * BNM);
  w.gmu1i = take(BN * d->Dx); w.gxn = take(BN * d->Dx); w.ginA = take(BN * d->Dx); w.ginB = take(BN * d->Dx);
  w.gm2_all = take(BT * d->Dx); w.gmu_init_all = take(BT * d->Dx); w.gmu0 = take((size_t)d->B * d->Dx);
  w.stats = take(FS_SLOTS * 32);
  w.gfeat_tmp = take(BT * (size_t)(d->Dfeat > 0 ? d->Dfeat : 1));
  w.gx0_tmp = take((size_t)d->B * (d->Dx0 > 0 ? d->Dx0 : 1));
  w.bytes = off;
  return w;
}

}  // namespace psvo

using namespace psvo;

extern "C" {

PSVO_API size_t psvo_ffbsi_backward_workspace_bytes(const psvo_dims* d, const psvo_params* p) {
  (void)p;
  if (d == nullptr) return 0;
  return carve_fb(d, nullptr).bytes;
}

PSVO_API int psvo_ffbsi_backward(const psvo_dims* d, const psvo_params* p, const float* obs, const float* bs_feat,
                                 const float* pre_x0, const float* Xs, const float* log_Ws, const float* lse,
                                 const float* eps_bw, uint64_t seed, const float* bw_Xs, const int32_t* pick,
                                 const float* f_lp, const float* g_lp, const float* Omega_or_bw_log_W,
                                 const float* bw_X_anc, const int32_t* anc2, const float* g_logZ_b, psvo_grads* grads,
                                 float* g_bs_feat, float* g_pre_x0, float* g_Xs, float* g_log_Ws, void* ws, size_t ws_bytes,
                                 psvo_stream_t stream) {
  if (int e = check_dims(d)) return e;
  PSVO_REQUIRE(p != nullptr && grads != nullptr, "params / grads is null");
  const bool boot = d->flags & PSVO_F_BOOTSTRAP, two_q = d->flags & PSVO_F_TWO_Q, wr = d->flags & PSVO_F_WITH_RESAMPLING;
  PSVO_REQUIRE(d->M >= 2 && d->M <= 256, "n_particles_for_BSim_proposal=%d must be in [2,256]", d->M);
  PSVO_REQUIRE(d->T >= 2, "backward simulation needs time >= 2");
  PSVO_REQUIRE(d->Dfeat >= 1, "Dfeat must be positive");
  const psvo_dist& fd = boot ? p->q1 : p->f;
  const psvo_dist& init = (boot && two_q) ? p->q0 : fd;
  psvo_dist_grad& gfd = boot ? grads->q1 : grads->f;
  psvo_dist_grad& ginit = (boot && two_q) ? grads->q0 : gfd;
  if (int e = check_mlp("f", fd.net, d->Dx, d->Dx)) return e;
  if (int e = check_mlp("g", p->g.net, d->Dx, d->Dy)) return e;
  if (int e = check_mlp("BSim_q_init", p->bsim_q_init.net, d->Dfeat, d->Dx)) return e;
  if (int e = check_mlp("q1_inv", p->q1_inv.net, d->Dx, d->Dx)) return e;
  if (int e = check_mlp("BSim_q2", p->bsim_q2.net, d->Dfeat, d->Dx)) return e;
  if (int e = check_mlp("init dist", init.net, d->Dx0, d->Dx)) return e;
  PSVO_REQUIRE(fd.net.n_hidden <= 1 && p->g.net.n_hidden <= 1 && p->q1_inv.net.n_hidden <= 1,
               "psvo_ffbsi_backward: per-particle networks with more than one hidden layer are not supported yet");
  PSVO_REQUIRE(obs && bs_feat && pre_x0 && Xs && log_Ws && lse && bw_Xs && pick && Omega_or_bw_log_W && g_logZ_b && g_Xs &&
               g_log_Ws && ws, "psvo_ffbsi_backward: null buffer");
  PSVO_REQUIRE(wr ? (bw_X_anc && anc2) : (f_lp && g_lp), "psvo_ffbsi_backward: missing %s forward outputs", wr ? "PSVOwR" : "PSVO");
  const FbWs w = carve_fb(d, static_cast<char*>(ws));
  PSVO_REQUIRE(ws_bytes >= w.bytes, "workspace too small: %zu < %zu", ws_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, T = d->T, N = d->N, M = d->M, Dx = d->Dx;
  const size_t BT = (size_t)B * T, BN = (size_t)B * N;
  const int DXP = pad_dx(Dx), DYP = pad_dy(d->Dy);

  // per-sequence means (as in the forward pass), cleared accumulators, coefficients
  if (int e = launch_sigma_block(sigma_src(d, p, true), w.sig, st)) return e;
  if (int e = launch_mlp_rows_forward(p->bsim_q2.net, bs_feat, (int64_t)BT, w.m2_all, st)) return e;
  if (int e = launch_mlp_rows_forward(p->bsim_q_init.net, bs_feat, (int64_t)BT, w.mu_init_all, st)) return e;
  if (int e = launch_mlp_rows_forward(init.net, pre_x0, B, w.mu0i, st)) return e;
  PSVO_CUDA(cudaMemsetAsync(w.gm2_all, 0, BT * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.gmu_init_all, 0, BT * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.gmu0, 0, (size_t)B * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.stats, 0, FS_SLOTS * 32 * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(w.ginA, 0, BN * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(g_Xs, 0, BT * N * Dx * sizeof(float), st));
  PSVO_CUDA(cudaMemsetAsync(g_log_Ws, 0, BT * N * sizeof(float), st));
  {
    ProfScope ps_(PK_FFBSI, st);
    ffbsi_coef_kernel<<<wr ? B * T : B, 256, 0, st>>>(f_lp, g_lp, Omega_or_bw_log_W, g_logZ_b, N, T, wr ? 1 : 0, w.coef);
    PSVO_LAUNCH_CHECK();
  }

  FbArgs a;
  a.f = fd.net; a.g = p->g.net; a.q1i = p->q1_inv.net;
  a.gf = gfd.net; a.gg = grads->g.net; a.gq1i = grads->q1_inv.net;
  a.sig = w.sig; a.obs = obs; a.Xs = Xs; a.log_Ws = log_Ws; a.lse = lse;
  a.mu_f = w.mu_f; a.mu1i = w.mu1i; a.m2_all = w.m2_all; a.mu_init_all = w.mu_init_all; a.mu0i = w.mu0i;
  a.X_next = wr ? bw_X_anc : bw_Xs;
  a.eps_bw = eps_bw; a.pick = pick; a.anc2 = anc2; a.coef = w.coef;
  a.candX = w.candX; a.candK = w.candK; a.candP = w.candP; a.gmu1i = w.gmu1i; a.gxn = w.gxn;
  a.gm2_all = w.gm2_all; a.gmu_init_all = w.gmu_init_all; a.gmu0 = w.gmu0; a.g_Xs = g_Xs; a.g_log_Ws = g_log_Ws;
  a.stats = w.stats; a.seed = seed;
  a.N = N; a.M = M; a.B = B; a.T = T; a.Dx = Dx; a.Dy = d->Dy; a.seq_offset = d->seq_offset; a.flags = d->flags;

  // candidate kernel: block = IL trajectories x M candidates, as large as shared memory allows
  const size_t fixedA = 5 * DXP + 32 + 4 + 64 + FS_SLOTS * 32 + 2 * (staged_fl(a.f, DXP, DXP) + staged_fl(a.g, DXP, DYP)) +
                        kBwdPairChunk * DXP + kBwdPairChunk;
  int blockA = 0; size_t smemA = 0;
  for (int target = 256; target >= 32; target /= 2) {
    int IL = target / M; if (IL < 1) IL = 1; if (IL > N) IL = N;
    const int blk = ((IL * M + 31) / 32) * 32;
    const int dmax = DXP > DYP ? DXP : DYP;
    const size_t fl = fixedA + blk + (size_t)blk * 2 * DXP + (size_t)2 * kBwdChunk * (blk + 1) + (size_t)blk * DXP + (size_t)blk * dmax;
    if (blk <= 256 && fl * sizeof(float) <= 224 * 1024) { blockA = blk; smemA = fl * sizeof(float); a.IL = IL; break; }
    if (IL == 1) break;
  }
  PSVO_REQUIRE(blockA > 0, "psvo_ffbsi_backward: networks / M=%d too large for one SM's shared memory", M);
  const int64_t workA = (int64_t)B * ((N + a.IL - 1) / a.IL);
  const int gridA = (int)(workA < num_sms() ? workA : num_sms());

  auto pick_block = [&](size_t fixed_floats, int* block, size_t* smem) -> int {
    int P = 256;
    while (P > 32 && N <= P / 2) P /= 2;
    for (;; P /= 2) {
      const size_t fl = fixed_floats + (size_t)2 * kBwdChunk * (P + 1) + (size_t)2 * P * DXP;
      if (fl * sizeof(float) <= 220 * 1024 || P == 32) { *block = P; *smem = fl * sizeof(float); break; }
    }
    PSVO_REQUIRE(*smem <= 227 * 1024, "psvo_ffbsi_backward: networks too large for shared memory (%zu B)", *smem);
    return 0;
  };
  int blockB, blockC; size_t smemB, smemC;
  if (int e = pick_block(DXP + 2 * staged_fl(a.f, DXP, DXP) + kBwdPairChunk * (DXP + 2), &blockB, &smemB)) return e;
  if (N <= 64) {
    // few particles: a 256-thread block whose extra threads split every particle's candidate loop (ffbsi_bwd_pair_kernel)
    const size_t fl = DXP + 2 * staged_fl(a.f, DXP, DXP) + kBwdPairChunk * (DXP + 2) + (size_t)2 * kBwdChunk * (256 + 1) + (size_t)2 * 256 * DXP;
    if (fl * sizeof(float) <= 220 * 1024) { blockB = 256; smemB = fl * sizeof(float); }
  }
  if (int e = pick_block(2 * staged_fl(a.q1i, DXP, DXP), &blockC, &smemC)) return e;
  auto grid_of = [&](int block) {
    const int64_t work = (int64_t)B * ((N + block - 1) / block);
    return (int)(work < num_sms() ? work : num_sms());
  };

  float* gin = w.ginA;
  float* gin_next = w.ginB;
  const bool pair_deferred = w.cand_steps == T && T > 1 && w.mu_f_all != nullptr;
  const size_t BNM_ = BN * (size_t)M;
  for (int t = 0; t < T; ++t) {
    a.t = t; a.gin = gin; a.gin_next = gin_next;
    if (pair_deferred) { a.candX = w.candX + (size_t)t * BNM_ * DXP; a.candK = w.candK + (size_t)t * BNM_; a.candP = w.candP + (size_t)t * BNM_; }
    int rc = 0;
    if (w.mu_f_all != nullptr) {
      // batch 1: both inputs (the filter's particles, the simulated trajectories) are known for every step: row (t, n) of the
      // input -> row (t, n) of the all-steps arrays, computed once before the loop ("sequences" = time steps)
      if (t == 0) {
        if ((rc = launch_particle_mlp_dx(fd.net, Xs, (long long)N * Dx, T, N, Dx, w.mu_f_all, st))) return rc;
        if ((rc = launch_particle_mlp_dx(p->q1_inv.net, a.X_next, (long long)N * Dx, T, N, Dx, w.mu1i_all, st))) return rc;
      }
      a.mu_f = t >= 1 ? w.mu_f_all + (size_t)(t - 1) * N * Dx : w.mu_f;
      a.mu1i = t <= T - 2 ? w.mu1i_all + (size_t)(t + 1) * N * Dx : w.mu1i;
    } else {
      if (t >= 1) if ((rc = launch_particle_mlp_dx(fd.net, Xs + (size_t)(t - 1) * N * Dx, (long long)T * N * Dx, B, N, Dx, w.mu_f, st))) return rc;
      if (t <= T - 2) if ((rc = launch_particle_mlp_dx(p->q1_inv.net, a.X_next + (size_t)(t + 1) * N * Dx, (long long)T * N * Dx, B, N, Dx, w.mu1i, st))) return rc;
    }
    rc = -1;
#define PSVO_CASE_A(X, Y) if (DXP == X && DYP == Y) rc = launch_fb(ffbsi_bwd_cand_kernel<X, Y>, a, gridA, blockA, smemA, st);
    PSVO_CASE_A(4, 4) PSVO_CASE_A(4, 8) PSVO_CASE_A(4, 16) PSVO_CASE_A(4, 32)
    PSVO_CASE_A(8, 4) PSVO_CASE_A(8, 8) PSVO_CASE_A(8, 16) PSVO_CASE_A(8, 32)
    PSVO_CASE_A(16, 4) PSVO_CASE_A(16, 8) PSVO_CASE_A(16, 16) PSVO_CASE_A(16, 32)
#undef PSVO_CASE_A
    if (rc) return rc;
    if (t >= 1 && !pair_deferred) {
      rc = DXP == 4 ? launch_fb(ffbsi_bwd_pair_kernel<4>, a, grid_of(blockB), blockB, smemB, st)
                    : (DXP == 8 ? launch_fb(ffbsi_bwd_pair_kernel<8>, a, grid_of(blockB), blockB, smemB, st)
                                : launch_fb(ffbsi_bwd_pair_kernel<16>, a, grid_of(blockB), blockB, smemB, st));
      if (rc) return rc;
    }
    if (t <= T - 2) {
      if (wr) PSVO_CUDA(cudaMemsetAsync(gin_next, 0, BN * Dx * sizeof(float), st));
      rc = DXP == 4 ? launch_fb(ffbsi_bwd_traj_kernel<4>, a, grid_of(blockC), blockC, smemC, st)
                    : (DXP == 8 ? launch_fb(ffbsi_bwd_traj_kernel<8>, a, grid_of(blockC), blockC, smemC, st)
                                : launch_fb(ffbsi_bwd_traj_kernel<16>, a, grid_of(blockC), blockC, smemC, st));
      if (rc) return rc;
    }
    float* tmp = gin; gin = gin_next; gin_next = tmp;
  }
  if (pair_deferred && T >= 2) {
    a.pair_all = 1; a.candX = w.candX; a.candK = w.candK; a.candP = w.candP; a.mu_f = w.mu_f_all;
    const long long gB = (long long)grid_of(blockB) * (T - 1);
    const int gridB = (int)(gB < 4LL * num_sms() ? gB : 4LL * num_sms());
    const int rc = DXP == 4 ? launch_fb(ffbsi_bwd_pair_kernel<4>, a, gridB, blockB, smemB, st)
                            : (DXP == 8 ? launch_fb(ffbsi_bwd_pair_kernel<8>, a, gridB, blockB, smemB, st)
                                        : launch_fb(ffbsi_bwd_pair_kernel<16>, a, gridB, blockB, smemB, st));
    if (rc) return rc;
    a.pair_all = 0;
  }

  // per-sequence networks: BSim_q2 and BSim_q_init over (b,t) rows, the initial density's network over b rows
  if (int e = launch_mlp_rows_backward(p->bsim_q2.net, bs_feat, w.gm2_all, (int64_t)BT, g_bs_feat, &grads->bsim_q2.net, st)) return e;
  if (int e = launch_mlp_rows_backward(p->bsim_q_init.net, bs_feat, w.gmu_init_all, (int64_t)BT, g_bs_feat ? w.gfeat_tmp : nullptr,
                                       &grads->bsim_q_init.net, st)) return e;
  if (g_bs_feat != nullptr) {
    ProfScope ps_(PK_FFBSI, st);
    fb_add_inplace_kernel<<<ceil_div((int64_t)BT * d->Dfeat, 256), 256, 0, st>>>(g_bs_feat, w.gfeat_tmp, (long long)BT * d->Dfeat);
    PSVO_LAUNCH_CHECK();
  }
  if (int e = launch_mlp_rows_backward(init.net, pre_x0, w.gmu0, B, g_pre_x0, &ginit.net, st)) return e;

  FbSigmaArgs sa;
  sa.stats = w.stats; sa.sig = w.sig; sa.Dx = Dx; sa.Dy = d->Dy; sa.flags = d->flags;
  const SigmaSrc src = sigma_src(d, p, true);
  for (int k = 0; k < SIG_SLOTS; ++k) { sa.con[k] = src.con[k]; sa.smin[k] = src.smin[k]; sa.gcon[k] = nullptr; }
  sa.gcon[SIG_Q0] = grads->q0.sigma_con; sa.gcon[SIG_Q1] = grads->q1.sigma_con; sa.gcon[SIG_F] = grads->f.sigma_con;
  sa.gcon[SIG_G] = grads->g.sigma_con; sa.gcon[SIG_BQI] = grads->bsim_q_init.sigma_con;
  sa.gcon[SIG_Q1I] = grads->q1_inv.sigma_con; sa.gcon[SIG_BQ2] = grads->bsim_q2.sigma_con;
  {
    ProfScope ps_(PK_FFBSI, st);
    ffbsi_sigma_grad_kernel<<<1, 32, 0, st>>>(sa);
    PSVO_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
