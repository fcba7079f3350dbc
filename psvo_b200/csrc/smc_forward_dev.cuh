// smc_forward_dev.cuh -- device code of the forward filter shared by the per-step kernels (smc_forward.cu), the persistent
// whole-filter kernel (smc_persistent.cu) and the thread-block-cluster whole-filter kernel (smc_cluster.cu): one time step of
// one tile of particles (propagate_tile) and normalise + resample + gather of one sequence (resample_body, gather_children).
// All three run this very code, so their outputs agree bit for bit.
//
// Reference: SVO.SMC (src/SMC/SVO.py:60-180), sample_from_2_dist diag branch (:182-197,228-232), resample_X /
// get_resample_idx (:243-300), compute_log_ZSMC (:302-311).
#pragma once
#include <stdlib.h>
#include "mlp_tile.cuh"
#include "smc_kernels.h"
#include "resample_group.cuh"

namespace psvo {

// ------------------------------------------------------------------------------------------------------
// K1: propagate + weight
// ------------------------------------------------------------------------------------------------------
template <int DXP>
struct StepConsts {        // per-CTA constants in shared memory
  float s1[DXP];           // stddev of d1 (q0 at t = 0, q1 afterwards)
  float sf[DXP];           // stddev used by f_lp
  float sp[DXP];           // proposal stddev (fused or s1)
  float inv1[DXP], inv2[DXP];
  float inv_sp[DXP], inv_sf[DXP];   // reciprocals: the standardised residuals are multiplications in the particle loop
  float sg[PSVO_MAX_DY], inv_sg[PSVO_MAX_DY];
  float ld_p, ld_f, ld_g;  // sum log stddev
};

// per-step constants of one CTA (t = 0 uses q0's stddev as d1, later steps q1's)
template <int DXP>
__device__ __forceinline__ void step_consts_init(const PropArgs& a, bool t0, StepConsts<DXP>* C) {
  const int Dx = a.Dx, Dy = a.Dy;
  const bool two_q = a.flags & PSVO_F_TWO_Q, boot = a.flags & PSVO_F_BOOTSTRAP;
  const bool poisson = a.flags & PSVO_F_POISSON;
  const bool f_is_d1 = boot && two_q;
  if (threadIdx.x < DXP) {
    const int d = threadIdx.x;
    const float s1 = a.sig[(t0 ? SIG_Q0 : SIG_Q1) * 32 + d];
    const float s2 = a.sig[SIG_Q2 * 32 + d];
    const float sfd = f_is_d1 ? s1 : a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
    const float i1 = 1.0f / s1, i2 = 1.0f / s2;
    C->s1[d] = s1; C->sf[d] = sfd; C->inv1[d] = i1; C->inv2[d] = i2;
    const float spv = two_q ? 1.0f / (i1 + i2) : s1;
    C->sp[d] = spv; C->inv_sp[d] = 1.0f / spv; C->inv_sf[d] = 1.0f / sfd;
  }
  for (int i = threadIdx.x; i < PSVO_MAX_DY; i += blockDim.x) { const float v = a.sig[SIG_G * 32 + i]; C->sg[i] = v; C->inv_sg[i] = 1.0f / v; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float lp = 0.f, lf = 0.f, lg = 0.f;
    for (int d = 0; d < Dx; ++d) { lp += logf(C->sp[d]); lf += logf(C->sf[d]); }
    if (!poisson) for (int d = 0; d < Dy; ++d) lg += logf(C->sg[d]);
    C->ld_p = lp; C->ld_f = lf; C->ld_g = lg;
  }
  __syncthreads();
}

// one tile of blockDim.x * PPT particles of sequence b at time t: propagate + weight  (shared by the per-step kernel
// and the persistent whole-filter kernel, so both produce the same bits)
template <int DXP, int DYP, int PPT>
__device__ __forceinline__ void propagate_tile(const PropArgs& a, const StepConsts<DXP>* C, const MlpS<DXP, DXP>& q1s,
                                               const MlpS<DXP, DXP>& fs, const MlpS<DXP, DYP>& gs, float* scratch,
                                               int b, int n0, int t, int n_end = 0x7fffffff) {
  const int Dx = a.Dx, Dy = a.Dy, N = a.N;
  n_end = n_end < N ? n_end : N;                 // particles [n0, n_end) of the tile are live (cluster kernel: a CTA's own range)
  const bool two_q = a.flags & PSVO_F_TWO_Q, boot = a.flags & PSVO_F_BOOTSTRAP;
  const bool poisson = a.flags & PSVO_F_POISSON;
  const bool t0 = t == 0;
  const bool f_is_d1 = boot && two_q;            // f_lp = d1_mvn.log_prob(X)  (SVO.py:230)
  const bool sep_f = !boot;                      // a distinct f network exists
  const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;
  const float neg_logN = -logf((float)N);
  {
    const size_t bt = (size_t)b * a.T + t;

    float x[PPT][DXP];        // parent, later the new particle
    float mu1[PPT][DXP];
    int n_of[PPT];
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
      n_of[pp] = n0 + pp * blockDim.x + threadIdx.x;
      const bool live = n_of[pp] < n_end;
      if (!t0) {
        const float* src = a.X_anc + ((bt - 1) * N + (live ? n_of[pp] : 0)) * Dx;
        if (Dx == DXP) {
#pragma unroll
          for (int v = 0; v < DXP / 4; ++v) {
            const float4 q = reinterpret_cast<const float4*>(src)[v];
            x[pp][4 * v] = q.x; x[pp][4 * v + 1] = q.y; x[pp][4 * v + 2] = q.z; x[pp][4 * v + 3] = q.w;
          }
        } else {
#pragma unroll
          for (int d = 0; d < DXP; ++d) x[pp][d] = d < Dx ? src[d] : 0.0f;
        }
      }
    }

    // ---- d1 mean: MLP q1 on the parent (t >= 1) or the per-sequence q0 mean (t = 0)
    float muf[PPT][DXP];
    if (!t0) {
      mlp_eval<DXP, DXP, PPT>(q1s, x, mu1, scratch, a.scratch_half);
      if (sep_f) mlp_eval<DXP, DXP, PPT>(fs, x, muf, scratch, a.scratch_half);
    } else {
#pragma unroll
      for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
        for (int d = 0; d < DXP; ++d) {
          mu1[pp][d] = d < Dx ? a.mu0[(size_t)b * Dx + d] : 0.0f;
          if (!f_is_d1) muf[pp][d] = d < Dx ? a.muf0[(size_t)b * Dx + d] : 0.0f;
        }
    }
    if (f_is_d1 || (!t0 && !sep_f)) {
#pragma unroll
      for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
        for (int d = 0; d < DXP; ++d) muf[pp][d] = mu1[pp][d];
    }

    // ---- proposal, sample, q / f log densities
    float flp[PPT], qlp[PPT];
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
      const bool live = n_of[pp] < n_end;
      const size_t row = bt * N + (live ? n_of[pp] : 0);
      float e[DXP];
      if (a.eps != nullptr) {
        const float* src = a.eps + row * Dx;
        if (Dx == DXP) {
#pragma unroll
          for (int v = 0; v < DXP / 4; ++v) {
            const float4 q = reinterpret_cast<const float4*>(src)[v];
            e[4 * v] = q.x; e[4 * v + 1] = q.y; e[4 * v + 2] = q.z; e[4 * v + 3] = q.w;
          }
        } else {
#pragma unroll
          for (int d = 0; d < DXP; ++d) e[d] = d < Dx ? src[d] : 0.0f;
        }
      } else {
        const uint64_t gidx = ((uint64_t)(a.seq_offset + b) * a.T + t) * N + (live ? n_of[pp] : 0);
#pragma unroll
        for (int c = 0; c < DXP / 4; ++c) {
          if (4 * c < Dx) normal4(philox4x32_10(a.seed, gidx, c, 0u), &e[4 * c]);
          else { e[4 * c] = e[4 * c + 1] = e[4 * c + 2] = e[4 * c + 3] = 0.f; }
        }
      }
      float qs = 0.f, fsum = 0.f;
#pragma unroll
      for (int d = 0; d < DXP; ++d) {
        float m = mu1[pp][d];
        if (two_q) {
          const float m2 = d < Dx ? a.mu2_all[bt * Dx + d] : 0.0f;
          m = C->sp[d] * (C->inv1[d] * mu1[pp][d] + C->inv2[d] * m2);
        }
        const float X = m + C->sp[d] * e[d];
        const float zq = (X - m) * C->inv_sp[d];
        const float zf = (X - muf[pp][d]) * C->inv_sf[d];
        if (d < Dx) { qs = fmaf(zq, zq, qs); fsum = fmaf(zf, zf, fsum); }
        x[pp][d] = d < Dx ? X : 0.0f;
      }
      qlp[pp] = -0.5f * qs - C->ld_p - cx;
      flp[pp] = -0.5f * fsum - C->ld_f - cx;
      if (live) {
        float* dst = a.Xs + row * Dx;
        if (Dx == DXP) {
#pragma unroll
          for (int v = 0; v < DXP / 4; ++v)
            reinterpret_cast<float4*>(dst)[v] = make_float4(x[pp][4 * v], x[pp][4 * v + 1], x[pp][4 * v + 2], x[pp][4 * v + 3]);
        } else {
#pragma unroll
          for (int d = 0; d < DXP; ++d) if (d < Dx) dst[d] = x[pp][d];
        }
      }
    }

    // ---- emission
    float mug[PPT][DYP];
    mlp_eval<DXP, DYP, PPT>(gs, x, mug, scratch, a.scratch_half);
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
      const bool live = n_of[pp] < n_end;
      float gsum = 0.f;
#pragma unroll
      for (int d = 0; d < DYP; ++d) {
        if (d < Dy) {
          const float y = a.obs[bt * Dy + d];
          float z;
          if (poisson) z = y - (softplus(mug[pp][d]) + 1e-6f);
          else z = (y - mug[pp][d]) * C->inv_sg[d];
          gsum = fmaf(z, z, gsum);
        }
      }
      const float g_lp = -0.5f * gsum - C->ld_g - cy;
      // log_alpha = f + g - q ;  log_W = log_alpha + log Wbar_{t-1}   (SVO.py:97-98,151-152)
      float lwb = neg_logN;
      if (!t0 && a.logWbar != nullptr && live) lwb = a.logWbar[(size_t)b * N + n_of[pp]];
      if (live) a.log_Ws[bt * N + n_of[pp]] = ((flp[pp] + g_lp) - qlp[pp]) + lwb;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// K2: normalise + resample + gather.  gridDim = (B, parts): every CTA of a sequence rebuilds the sequence's CDF
// (N deterministic exps + one integer block scan -- cheap) and serves the children [part*chunk, (part+1)*chunk), so the
// dependent search + gather chains of one sequence spread over several SMs when B alone cannot fill the machine.
// shared memory: N uint64 (inclusive fixed-point CDF) + N floats (log weights) + chunk ints (ancestors of this part)
// ------------------------------------------------------------------------------------------------------
// normalise + resample + gather for sequence b at time t, children of part `part` of `nparts`; NC: X_t may be read
// through the non-coherent path (only when a previous kernel wrote it)
// sG (optional, K+1 ints with K = 2^kbits >= N): guide table G[k] = #{j : C[j] <= B_k} over the K equal slices of the
// uniform's 24-bit range (B_k = draw_threshold(total, k 2^24/K)).  A child whose uniform falls in slice k has its ancestor in
// [G[k], G[k+1]] -- about one particle wide -- so the inverse-CDF search is 1-2 probes of the CDF instead of log2 N
// bank-conflicting ones; the result is the same upper bound, bit for bit.
template <bool NC>
__device__ __forceinline__ void gather_children(const ResampleArgs& a, size_t bt, const int* sAnc, int n_lo, int n_hi, int tid, int nt);

template <bool NC>
__device__ __forceinline__ void resample_body(const ResampleArgs& a, int b, int part, int nparts, int t, unsigned char* smraw,
                                              unsigned long long* sscr, float* fscr, int* sG = nullptr, int kbits = 0) {
  const int N = a.N;
  unsigned long long* sC = reinterpret_cast<unsigned long long*>(smraw);
  float* sLW = reinterpret_cast<float*>(sC + N);
  int* sAnc = reinterpret_cast<int*>(sLW + N);
  const int chunk = (N + nparts - 1) / nparts;
  const int n_lo = min(part * chunk, N), n_hi = min(n_lo + chunk, N);
  const size_t bt = (size_t)b * a.T + t;
  const float* lw = a.log_Ws + bt * N;
  const int tid = threadIdx.x, nt = blockDim.x;

  float mx = -INFINITY;
  for (int i = tid; i < N; i += nt) { const float v = lw[i]; sLW[i] = v; mx = fmaxf(mx, v); }
  mx = block_max(mx, fscr);

  // integer CDF: every warp owns a contiguous run of the particles and walks it 32 at a time (lane = particle:
  // conflict-free shared-memory rows), warp-shuffle inclusive scan + running carry; then the warp totals are scanned
  // and added back.  Integer addition is associative, so the CDF does not depend on this decomposition.
  const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
  const int per_warp = (((N + nw - 1) / nw) + 31) & ~31;
  const int w_lo = min(wid * per_warp, N), w_hi = min(w_lo + per_warp, N);
  unsigned long long carry = 0ull;
  for (int base = w_lo; base < w_hi; base += 32) {
    const int i = base + lane;
    unsigned long long v = i < w_hi ? fixed_weight(sLW[i], mx) : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long nb = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += nb;
    }
    v += carry;
    if (i < w_hi) sC[i] = v;
    carry = __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) sscr[wid] = carry;
  __syncthreads();
  if (wid == 0) {
    const unsigned long long wt = lane < nw ? sscr[lane] : 0ull;
    unsigned long long inc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long nb = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += nb;
    }
    if (lane < nw) sscr[lane] = inc - wt;           // exclusive warp offsets
    if (lane == 31) sscr[32] = inc;                  // grand total
  }
  __syncthreads();
  const unsigned long long total = sscr[32];
  const unsigned long long off = sscr[wid];
  if (off != 0ull) for (int i = w_lo + lane; i < w_hi; i += 32) sC[i] += off;
  __syncthreads();

  // logsumexp from the same integer total: deterministic and independent of the thread count
  const float lse = mx + (logf((float)total) - 32.0f * kLn2);
  if (tid == 0 && part == 0) {
    if (a.lse != nullptr) a.lse[bt] = lse;
    if (a.logZ_b != nullptr) a.logZ_b[b] = (t == 0 ? 0.0f : a.logZ_b[b]) + lse;
  }

  const bool resample = a.flags & PSVO_F_RESAMPLE;
  int top = 1;                                   // largest power of two <= N
  while (top * 2 <= N) top *= 2;
  const bool guided = sG != nullptr && resample && a.forced_anc == nullptr;
  const int K = 1 << kbits, gs = 24 - kbits;
  if (guided) {
    for (int i = tid; i <= K; i += nt) sG[i] = 0;
    __syncthreads();
    // kf(j) = min{k : B_k >= C[j]}: float estimate, exact integer fix-up; histogram, then G = inclusive scan
    const float scale = (float)K / (float)total;
    for (int j = tid; j < N; j += nt) {
      const unsigned long long c = sC[j];
      int kf = (int)((float)c * scale);
      kf = kf < 0 ? 0 : (kf > K ? K : kf);
      while (kf > 0 && draw_threshold(total, (uint32_t)(kf - 1) << gs) >= c) --kf;
      while (kf < K && draw_threshold(total, (uint32_t)kf << gs) < c) ++kf;
      atomicAdd(&sG[kf], 1);
    }
    __syncthreads();
    const int nG = K + 1;
    const int pw = (((nG + nw - 1) / nw) + 31) & ~31;
    const int g_lo = min(wid * pw, nG), g_hi = min(g_lo + pw, nG);
    int gcarry = 0;
    for (int base = g_lo; base < g_hi; base += 32) {
      const int i = base + lane;
      int v = i < g_hi ? sG[i] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += nb;
      }
      v += gcarry;
      if (i < g_hi) sG[i] = v;
      gcarry = __shfl_sync(0xffffffffu, v, 31);
    }
    int* iscr = reinterpret_cast<int*>(fscr);      // 33 words, free after block_max
    if (lane == 0) iscr[wid] = gcarry;
    __syncthreads();
    if (wid == 0) {
      const int wt = lane < nw ? iscr[lane] : 0;
      int inc = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += nb;
      }
      if (lane < nw) iscr[lane] = inc - wt;
    }
    __syncthreads();
    const int goff = iscr[wid];
    if (goff != 0) for (int i = g_lo + lane; i < g_hi; i += 32) sG[i] += goff;
    __syncthreads();
  }
  // four CONSECUTIVE children per thread at a time: their uniforms are the four words of one Philox block (uniform_bits), and
  // their branch-free binary searches interleave, hiding the shared-memory latency
  const bool systematic = a.flags & PSVO_F_SYSTEMATIC;
  const uint64_t gbase = ((uint64_t)(a.seq_offset + b) * a.T + t) * N;
  for (int n0 = n_lo + 4 * tid; n0 < n_hi; n0 += 4 * nt) {
    unsigned long long thr[4];
    int anc[4];
    bool search[4];
    int kb[4] = {0, 0, 0, 0};
    const bool draw = resample && a.forced_anc == nullptr && a.u == nullptr;
    const bool block4 = draw && !systematic && ((gbase + (uint64_t)n0) & 3ull) == 0ull;
    Philox4 blk;
    blk.v[0] = blk.v[1] = blk.v[2] = blk.v[3] = 0u;
    if (block4) blk = philox4x32_10(a.seed, (gbase + (uint64_t)n0) >> 2, 0u, a.stream_id);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int n = n0 + k;
      anc[k] = n; search[k] = false; thr[k] = 0ull;
      if (n < n_hi && resample) {
        if (a.forced_anc != nullptr) {
          const int f = a.forced_anc[bt * N + n];
          anc[k] = f < 0 ? 0 : (f >= N ? N - 1 : f);
        } else {
          const int nu = systematic ? 0 : n;                          // systematic: one draw per (b, t), particle 0's
          uint32_t u24;
          if (a.u != nullptr) u24 = u24_from_float(a.u[bt * N + nu]);
          else if (block4) u24 = blk.v[k] >> 8;
          else u24 = uniform_bits(a.seed, gbase + (uint64_t)nu, a.stream_id) >> 8;
          if (systematic) u24 = (uint32_t)((((uint64_t)n << 24) + u24) / (uint64_t)N);   // floor(2^24 (n + u_0) / N)
          thr[k] = draw_threshold(total, u24);
          search[k] = true;
          kb[k] = min((int)(u24 >> gs), K - 1);
        }
      }
    }
    // #{j : C[j] <= thr} (upper bound) by descending powers of two; identical to upper_bound_u64 on a monotone CDF
    int pos[4] = {0, 0, 0, 0};
    if (guided) {
      int hi[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { pos[k] = search[k] ? sG[kb[k]] : 0; hi[k] = search[k] ? sG[kb[k] + 1] : 0; }
      bool any = true;
      while (any) {                              // the four bracketed searches advance together
        any = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (pos[k] < hi[k]) {
            const int mid = (pos[k] + hi[k]) >> 1;
            if (sC[mid] <= thr[k]) pos[k] = mid + 1; else hi[k] = mid;
            any = true;
          }
        }
      }
    } else {
      for (int s = top; s > 0; s >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int cand = pos[k] + s;
          if (search[k] && cand <= N && sC[cand - 1] <= thr[k]) pos[k] = cand;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int n = n0 + k;
      if (n >= n_hi) continue;
      if (search[k]) anc[k] = pos[k] < N ? pos[k] : N - 1;
      if (!resample) a.logWbar[(size_t)b * N + n] = sLW[n] - lse;      // SVO.py:102,156 (IWAE keeps it differentiable)
      a.anc[bt * N + n] = anc[k];
      sAnc[n - n_lo] = anc[k];
    }
  }
  __syncthreads();

  if (a.skip_gather) return;                     // the propagate launch of t+1 gathers (PropArgs::anc_prev)
  gather_children<NC>(a, bt, sAnc, n_lo, n_hi, tid, nt);
}

// ancestor gather X_anc[t][n] = X_t[anc[n]] for the children [n_lo, n_hi) (sAnc[n - n_lo] = ancestor of child n): consecutive
// threads write consecutive 16-byte pieces of the children
template <bool NC>
__device__ __forceinline__ void gather_children(const ResampleArgs& a, size_t bt, const int* sAnc, int n_lo, int n_hi, int tid, int nt) {
  const int N = a.N, Dx = a.Dx;
  const float* Xt = a.Xs + bt * N * Dx;
  float* Xa = a.X_anc + bt * N * Dx;
  const int nchild = n_hi - n_lo;
  if ((Dx & 3) == 0) {
    const int v4 = Dx >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(Xt);
    float4* dst4 = reinterpret_cast<float4*>(Xa) + (size_t)n_lo * v4;
    const int pieces = nchild * v4;
    // four independent row reads in flight per thread before the first store (the loop is latency-bound otherwise)
    for (int q0 = tid; q0 < pieces; q0 += 4 * nt) {
      float4 r[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = q0 + k * nt;
        if (q < pieces) {
          const int c = q / v4, v = q - c * v4;
          const float4* sp4 = src4 + (size_t)sAnc[c] * v4 + v;
          r[k] = NC ? __ldg(sp4) : *sp4;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int q = q0 + k * nt;
        if (q < pieces) dst4[q] = r[k];
      }
    }
  } else {
    const int pieces = nchild * Dx;
    float* dst = Xa + (size_t)n_lo * Dx;
    for (int q = tid; q < pieces; q += nt) {
      const int c = q / Dx, d = q - c * Dx;
      dst[q] = Xt[(size_t)sAnc[c] * Dx + d];
    }
  }
}

}  // namespace psvo
