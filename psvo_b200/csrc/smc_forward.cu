// smc_forward.cu -- forward particle filter: propagate + weight (K1) and normalise + resample + gather (K2).
//
// Reference: SVO.SMC (src/SMC/SVO.py:60-180), sample_from_2_dist diag branch (:182-197,228-232),
// resample_X / get_resample_idx (:243-300), compute_log_ZSMC (:302-311).  SURVEY.md App. A.3-A.5.
//
// Per time step two kernels run back to back on the caller's stream:
//   smc_propagate : grid over (sequence, particle tile).  One thread owns PPT particles: loads the resampled
//                   parent X_anc[t-1], evaluates MLP q1 (and f), fuses with the per-sequence q2 mean, samples
//                   X_t, evaluates MLP g and the three log densities, writes X_t and log_W_t.  Weights live in
//                   shared memory and are read as warp broadcasts; nothing is exchanged between threads.
//   smc_resample  : one CTA per sequence.  log-sum-exp of log_W_t, order-independent fixed-point CDF (integer
//                   block scan), inverse-CDF binary search per child, ancestor gather X_anc[t] = X_t[a].
// HBM layout is sequence-major (B,T,N,D): every (b,t) slab a kernel touches is contiguous.
#include <stdlib.h>
#include "mlp_tile.cuh"
#include <cstdlib>
#include "smc_kernels.h"
#include "resample_group.cuh"

namespace psvo {

// ------------------------------------------------------------------------------------------------------
// K0: sigma block.  sig[slot*32 + d] = max(softplus(sigma_con[d]), sigma_min)  (1 where the dist has none)
// ------------------------------------------------------------------------------------------------------
__global__ void sigma_block_kernel(SigmaSrc src, float* __restrict__ sig) {
  const int slot = blockIdx.x, d = threadIdx.x;
  if (slot >= SIG_SLOTS || d >= 32) return;
  float v = 1.0f;
  if (src.con[slot] != nullptr && d < src.dim[slot]) v = sigma_from_con(src.con[slot][d], src.smin[slot]);
  sig[slot * 32 + d] = v;
}

int launch_sigma_block(const SigmaSrc& src, float* sig, cudaStream_t st) {
  ProfScope ps_(PK_SIGMA, st);
  sigma_block_kernel<<<SIG_SLOTS, 32, 0, st>>>(src, sig);
  PSVO_LAUNCH_CHECK();
  return 0;
}

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

template <int DXP, int DYP, int PPT>
__global__ void __launch_bounds__(256) smc_propagate_kernel(PropArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int N = a.N;
  const bool t0 = a.t == 0;
  const bool sep_f = !(a.flags & PSVO_F_BOOTSTRAP);

  MlpS<DXP, DXP> q1s, fs;
  MlpS<DXP, DYP> gs;
  float* p = smem;
  StepConsts<DXP>* C = reinterpret_cast<StepConsts<DXP>*>(p);
  p += (sizeof(StepConsts<DXP>) / sizeof(float) + 3) & ~3;
  if (!t0) p = mlp_stage<DXP, DXP>(a.q1, p, q1s);
  if (!t0 && sep_f) p = mlp_stage<DXP, DXP>(a.f, p, fs);
  p = mlp_stage<DXP, DYP>(a.g, p, gs);
  float* scratch = p;
  step_consts_init<DXP>(a, t0, C);
  // programmatic dependent launch (PropArgs::pdl): the weights and constants above were staged while the resampling launch of
  // step t-1 was still running; its outputs (X_anc, log Wbar) are first read below
  griddep_launch_dependents();
  griddep_wait();

  const int tile = blockDim.x * PPT;
  const int tiles_per_seq = (N + tile - 1) / tile;
  const int work = a.B * tiles_per_seq;
  for (int wi = blockIdx.x; wi < work; wi += gridDim.x) {
    const int b = wi / tiles_per_seq;
    propagate_tile<DXP, DYP, PPT>(a, C, q1s, fs, gs, scratch, b, (wi - b * tiles_per_seq) * tile, a.t);
  }
}

template <int DXP, int DYP, int PPT>
static int launch_propagate_t(const PropArgs& a, int block, int grid, size_t smem, cudaStream_t st) {
  auto kern = smc_propagate_kernel<DXP, DYP, PPT>;
  PSVO_DYN_SMEM(kern, smem);
  ProfScope ps_(PK_PROP, st);
  PSVO_CUDA(launch_pdl(kern, dim3(grid), dim3(block), smem, st, a.pdl != 0, a));
  PSVO_LAUNCH_CHECK();
  return 0;
}

PropPlan plan_propagate(const PropArgs& a) {
  PropPlan pl;
  const int DXP = pad_dx(a.Dx), DYP = pad_dy(a.Dy);
  const bool sep_f = !(a.flags & PSVO_F_BOOTSTRAP);
  const bool deep = a.q1.n_hidden >= 2 || a.g.n_hidden >= 2 || (sep_f && a.f.n_hidden >= 2);
  const int64_t rows = (int64_t)a.B * a.N;
  // two particles per thread once there is enough work to fill the machine twice over
  pl.ppt = (!deep && rows >= (int64_t)2 * 256 * num_sms()) ? 2 : 1;
  pl.block = deep ? 128 : 256;
  while (pl.block > 32 && a.N <= pl.block / 2) pl.block /= 2;     // no idle warps on small particle counts
  if (a.N < pl.block * pl.ppt) pl.ppt = 1;
  size_t fl = 0;
  auto add = [&](int din_pad, int dout_pad, const psvo_mlp& net) {
    if (net.n_hidden == 0) return (size_t)(din_pad * dout_pad + dout_pad);
    size_t n = (size_t)net.width[0] * din_pad + round4(net.width[0]);
    for (int k = 1; k < net.n_hidden; ++k) n += round4(net.width[k - 1] * net.width[k]) + round4(net.width[k]);
    return n + (size_t)net.width[net.n_hidden - 1] * dout_pad + dout_pad;
  };
  fl += 192;   // StepConsts upper bound (floats)
  fl += add(DXP, DXP, a.q1);
  if (sep_f) fl += add(DXP, DXP, a.f);
  fl += add(DXP, DYP, a.g);
  pl.scratch_half = 0;
  if (deep) {
    int hm = mlp_max_width(a.q1);
    hm = hm > mlp_max_width(a.g) ? hm : mlp_max_width(a.g);
    if (sep_f) hm = hm > mlp_max_width(a.f) ? hm : mlp_max_width(a.f);
    pl.scratch_half = hm * pl.block * pl.ppt;
    fl += 2 * (size_t)pl.scratch_half;
  }
  pl.smem = fl * sizeof(float);
  const int tile = pl.block * pl.ppt;
  const int64_t work = (int64_t)a.B * ((a.N + tile - 1) / tile);
  const int per_sm = pl.smem > 100 * 1024 ? 1 : (pl.smem > 70 * 1024 ? 2 : 3);
  const int64_t cap = (int64_t)num_sms() * per_sm;
  pl.grid = (int)(work < cap ? work : cap);
  if (pl.grid < 1) pl.grid = 1;
  return pl;
}

int launch_propagate(PropArgs a, const PropPlan& pl, cudaStream_t st) {
  if (a.tc_image != nullptr && tc_propagate_supported(a)) return launch_propagate_tc(a, st);
  a.scratch_half = pl.scratch_half;
  const int DXP = pad_dx(a.Dx), DYP = pad_dy(a.Dy);
  PSVO_REQUIRE(pl.smem <= 227 * 1024, "MLP weights (%zu B) exceed one SM's shared memory", pl.smem);
#define PSVO_CASE(X, Y)                                                                          \
  if (DXP == X && DYP == Y) {                                                                    \
    return pl.ppt == 2 ? launch_propagate_t<X, Y, 2>(a, pl.block, pl.grid, pl.smem, st)          \
                       : launch_propagate_t<X, Y, 1>(a, pl.block, pl.grid, pl.smem, st);         \
  }
  PSVO_CASE(4, 4) PSVO_CASE(4, 8) PSVO_CASE(4, 16) PSVO_CASE(4, 32)
  PSVO_CASE(8, 4) PSVO_CASE(8, 8) PSVO_CASE(8, 16) PSVO_CASE(8, 32)
  PSVO_CASE(16, 4) PSVO_CASE(16, 8) PSVO_CASE(16, 16) PSVO_CASE(16, 32)
#undef PSVO_CASE
  return fail("unsupported padded dims Dx=%d Dy=%d", a.Dx, a.Dy);
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

__global__ void __launch_bounds__(1024) smc_resample_kernel(ResampleArgs a, int guide_off, int kbits) {
  extern __shared__ __align__(16) unsigned char smraw[];
  __shared__ unsigned long long sscr[33];
  __shared__ float fscr[33];
  int* sG = guide_off > 0 ? reinterpret_cast<int*>(smraw + guide_off) : nullptr;
  griddep_launch_dependents();               // the next propagate launch may stage its weights while this kernel runs
  griddep_wait();                            // log_W_t / X_t of the propagate launch before this one
  resample_body<true>(a, blockIdx.x, blockIdx.y, gridDim.y, a.t, smraw, sscr, fscr, sG, kbits);
}

// One CTA of 1024 threads per sequence, thread-contiguous decomposition (resample_group.cuh): every thread owns a contiguous
// run of N / 1024 particles, so the CDF and the guide table need one shuffle scan per WARP instead of one per 32 particles,
// and the cross-warp step is a second shuffle scan every warp does for itself -- about half the instructions of
// resample_body and six block barriers.  Same indices bit for bit (integer sums do not depend on their grouping).
__global__ void __launch_bounds__(1024) smc_resample_lean_kernel(ResampleArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  griddep_launch_dependents();               // the next propagate launch may stage its weights while this kernel runs
  griddep_wait();                            // log_W_t / X_t of the propagate launch before this one
  const int b = blockIdx.x, t = a.t, N = a.N, Dx = a.Dx;
  resample_lean<false>(a, b, t, smraw, threadIdx.x, 1024, RgBlockBar{});
  if (a.skip_gather) return;
  __syncthreads();                           // this block's ancestor indices are in global memory: read them back through L2
  const size_t bt = (size_t)b * a.T + t;
  const int* anc = a.anc + bt * N;
  const float* Xt = a.Xs + bt * N * Dx;
  float* Xa = a.X_anc + bt * N * Dx;
  if ((Dx & 3) == 0) {
    const int v4 = Dx >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(Xt);
    float4* dst4 = reinterpret_cast<float4*>(Xa);
    const int pieces = N * v4;
    for (int q = threadIdx.x; q < pieces; q += blockDim.x) {
      const int c = q / v4, v = q - c * v4;
      dst4[q] = __ldg(src4 + (size_t)__ldcg(anc + c) * v4 + v);
    }
  } else {
    for (int q = threadIdx.x; q < N * Dx; q += blockDim.x) {
      const int c = q / Dx, d = q - c * Dx;
      Xa[q] = Xt[(size_t)__ldcg(anc + c) * Dx + d];
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Persistent whole-filter kernel: ONE launch per forward pass, one CTA per sequence looping over time.
// For small particle counts / few sequences (BASELINE configs 1-4: B = 1 chains, the reference's own operating
// point n_particles = 16..64) the two launches per time step are pure latency; here the T dependent steps run
// back to back with __syncthreads between the propagate and the resample phase.  State goes through the output tensors
// themselves (L1/L2-resident for one sequence), the arithmetic is the very same device code as the per-step kernels.
// ------------------------------------------------------------------------------------------------------
// narrow state / observation dimensions leave enough registers for 1024 threads per CTA (one particle per thread up to
// N = 1024: the per-step critical path is then a handful of barriers)
__host__ __device__ constexpr int persist_max_block(int DXP, int DYP) { return (DXP <= 8 && DYP <= 8) ? 1024 : 256; }

constexpr int kPersistLeanMinN = 512;

template <int DXP, int DYP>
__global__ void __launch_bounds__(persist_max_block(DXP, DYP), 1) smc_filter_persistent_kernel(PropArgs pa, ResampleArgs ra, int prop_floats) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned long long sscr[33];
  __shared__ float fscr[33];
  const int N = pa.N;
  const bool sep_f = !(pa.flags & PSVO_F_BOOTSTRAP);
  MlpS<DXP, DXP> q1s, fs;
  MlpS<DXP, DYP> gs;
  float* p = smem;
  StepConsts<DXP>* C0 = reinterpret_cast<StepConsts<DXP>*>(p);
  p += (sizeof(StepConsts<DXP>) / sizeof(float) + 3) & ~3;
  StepConsts<DXP>* C1 = reinterpret_cast<StepConsts<DXP>*>(p);
  p += (sizeof(StepConsts<DXP>) / sizeof(float) + 3) & ~3;
  p = mlp_stage<DXP, DXP>(pa.q1, p, q1s);
  if (sep_f) p = mlp_stage<DXP, DXP>(pa.f, p, fs);
  p = mlp_stage<DXP, DYP>(pa.g, p, gs);
  float* scratch = p;
  unsigned char* rs_smem = reinterpret_cast<unsigned char*>(smem + prop_floats);
  int* sAnc = reinterpret_cast<int*>(rs_smem + ((rg_smem_bytes(N, blockDim.x) + 15) & ~(size_t)15));
  step_consts_init<DXP>(pa, true, C0);
  step_consts_init<DXP>(pa, false, C1);

  const int tile = blockDim.x;
  for (int b = blockIdx.x; b < pa.B; b += gridDim.x) {
    for (int t = 0; t < pa.T; ++t) {
      const StepConsts<DXP>* C = t == 0 ? C0 : C1;
      for (int n0 = 0; n0 < N; n0 += tile) propagate_tile<DXP, DYP, 1>(pa, C, q1s, fs, gs, scratch, b, n0, t);
      __syncthreads();                     // log_W_t / X_t of this sequence written (block-scope visibility)
      if (N >= kPersistLeanMinN) {
        // thread-contiguous resampling (resample_group.cuh): guide table, one shuffle scan per warp, same indices bit for bit
        resample_lean<false>(ra, b, t, rs_smem, threadIdx.x, blockDim.x, RgBlockBar{}, sAnc);
        __syncthreads();
        if (!ra.skip_gather) gather_children<false>(ra, (size_t)b * ra.T + t, sAnc, 0, N, threadIdx.x, blockDim.x);
      } else {
        resample_body<false>(ra, b, 0, 1, t, rs_smem, sscr, fscr);      // few particles: the plain search is a shorter chain than the guide table's phases
      }
      __syncthreads();                     // X_anc_t / logWbar written
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// Cluster whole-filter kernel: ONE launch per forward pass, one thread-block CLUSTER (2 / 4 / 8 CTAs on neighbouring SMs) per
// sequence.  For a few long chains of many particles (BASELINE config 4: N = 2048, B = 1, T = 1000) one CTA per sequence is
// too little parallelism and per-step launches are pure latency (two launches and two grid drains per step).  Here CTA r of
// the cluster owns the particles AND the children [r N / C, (r+1) N / C): it propagates its particles, the cluster barrier
// (release / acquire at cluster scope) makes every CTA's log-weights and states visible, every CTA rebuilds the sequence's CDF
// (N deterministic exps + an integer scan: cheap next to a launch) and resamples + gathers ITS children -- whose states it
// is the only reader of at the next step, so one cluster barrier per time step is all the cross-CTA synchronisation.
// Same device code as the per-step kernels (propagate_tile, resample_body with part = cluster rank): same bits.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

template <int DXP, int DYP>
__global__ void __launch_bounds__(256, 1) smc_filter_cluster_kernel(PropArgs pa, ResampleArgs ra, int prop_floats, int csize) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned long long sscr[33];
  __shared__ float fscr[33];
  const int N = pa.N;
  const bool sep_f = !(pa.flags & PSVO_F_BOOTSTRAP);
  MlpS<DXP, DXP> q1s, fs;
  MlpS<DXP, DYP> gs;
  float* p = smem;
  StepConsts<DXP>* C0 = reinterpret_cast<StepConsts<DXP>*>(p);
  p += (sizeof(StepConsts<DXP>) / sizeof(float) + 3) & ~3;
  StepConsts<DXP>* C1 = reinterpret_cast<StepConsts<DXP>*>(p);
  p += (sizeof(StepConsts<DXP>) / sizeof(float) + 3) & ~3;
  p = mlp_stage<DXP, DXP>(pa.q1, p, q1s);
  if (sep_f) p = mlp_stage<DXP, DXP>(pa.f, p, fs);
  p = mlp_stage<DXP, DYP>(pa.g, p, gs);
  float* scratch = p;
  unsigned char* rs_smem = reinterpret_cast<unsigned char*>(smem + prop_floats);
  step_consts_init<DXP>(pa, true, C0);
  step_consts_init<DXP>(pa, false, C1);

  const int rank = (int)cluster_ctarank();
  const int b = blockIdx.x / csize;                          // clusters are consecutive CTAs: cluster id = sequence
  const int chunk = (N + csize - 1) / csize;                 // the same split resample_body uses for its children
  const int n_lo = min(rank * chunk, N), n_hi = min(n_lo + chunk, N);
  const int tile = blockDim.x;
  for (int t = 0; t < pa.T; ++t) {
    const StepConsts<DXP>* C = t == 0 ? C0 : C1;
    for (int n0 = n_lo; n0 < n_hi; n0 += tile) propagate_tile<DXP, DYP, 1>(pa, C, q1s, fs, gs, scratch, b, n0, t, n_hi);
    cluster_sync_all();                                      // log_W_t / X_t of the whole sequence written and visible
    // (the guide-table routine of resample_group.cuh restricted to the CTA's children was tried here: 14.9 vs 12.95 ms per
    // config-4 pass -- every CTA would build the whole table for an eighth of the children)
    resample_body<false>(ra, b, rank, csize, t, rs_smem, sscr, fscr);
    __syncthreads();                                         // X_anc_t / logWbar of this CTA's own children written
  }
}

// resampling region of the persistent kernel: resample_lean's arrays + the ancestors for the gather (N >= kPersistLeanMinN), or
// resample_body's CDF | log weights | ancestors
static size_t persist_resample_bytes(int N, int block) {
  if (N < kPersistLeanMinN) return (size_t)N * 16;
  return ((rg_smem_bytes(N, block) + 15) & ~(size_t)15) + (size_t)N * sizeof(int);
}

static bool persistent_env() {
  static const int v = [] { const char* e = getenv("PSVO_PERSISTENT"); return e == nullptr ? -1 : (e[0] == '0' ? 0 : 1); }();
  return v != 0;
}

bool persistent_filter_supported(const PropArgs& a, const PropPlan& pl, size_t* smem_out, int* block_out) {
  if (!persistent_env()) return false;
  if (a.tc_image != nullptr && tc_propagate_supported(a)) return false;
  // one CTA walks a whole sequence: worth it while the per-step launches are latency-bound, i.e. while one step of one
  // sequence is a small amount of work (measured on B200: config 1 2.1x and config 3 1.5x faster than per-step launches,
  // config 4 -- N = 2048, H = 64 -- 2x slower, so it stays on the multi-CTA per-step path)
  if (a.N > 2048 || a.B > 2 * num_sms()) return false;
  {
    auto macs = [](const psvo_mlp& net) {
      long long m = 0; int in = net.d_in;
      for (int k = 0; k < net.n_hidden; ++k) { m += (long long)in * net.width[k]; in = net.width[k]; }
      return m + (long long)in * net.d_out;
    };
    long long instr = 300 + macs(a.q1) + macs(a.g) + ((a.flags & PSVO_F_BOOTSTRAP) ? 0 : macs(a.f));
    if ((long long)a.N * instr > 600000) return false;
  }
  const int DXP = pad_dx(a.Dx), DYP = pad_dy(a.Dy);
  int block = ((a.N < 1024 ? a.N : 1024) + 31) & ~31;
  if (block > persist_max_block(DXP, DYP)) block = persist_max_block(DXP, DYP);
  // keep every sequence's CTA co-resident (one wave): at 64 registers per thread an SM holds 1024 threads
  while (block > 128 && (long long)a.B * block > (long long)num_sms() * 1024) block /= 2;
  if (block < pl.block) block = pl.block < persist_max_block(DXP, DYP) ? pl.block : persist_max_block(DXP, DYP);
  // propagate region: two StepConsts + all three networks + deep-MLP scratch for `block` threads, one particle each
  const bool sep_f = !(a.flags & PSVO_F_BOOTSTRAP);
  auto add = [&](int din_pad, int dout_pad, const psvo_mlp& net) {
    if (net.n_hidden == 0) return (size_t)(din_pad * dout_pad + dout_pad);
    size_t n = (size_t)net.width[0] * din_pad + round4(net.width[0]);
    for (int k = 1; k < net.n_hidden; ++k) n += round4(net.width[k - 1] * net.width[k]) + round4(net.width[k]);
    return n + (size_t)net.width[net.n_hidden - 1] * dout_pad + dout_pad;
  };
  size_t fl = 384 + add(DXP, DXP, a.q1) + (sep_f ? add(DXP, DXP, a.f) : 0) + add(DXP, DYP, a.g);
  if (pl.scratch_half > 0) fl += 2 * (size_t)(pl.scratch_half / (pl.block * pl.ppt)) * block;
  fl = (fl + 3) & ~(size_t)3;
  const size_t total = fl * sizeof(float) + persist_resample_bytes(a.N, block);
  if (total > 100 * 1024) return false;
  *smem_out = total; *block_out = block;
  return true;
}

int launch_filter_persistent(PropArgs pa, const ResampleArgs& ra, const PropPlan& pl, cudaStream_t st) {
  size_t smem = 0; int block = 0;
  PSVO_REQUIRE(persistent_filter_supported(pa, pl, &smem, &block), "persistent filter kernel does not support this shape");
  const int prop_floats = (int)((smem - persist_resample_bytes(pa.N, block)) / sizeof(float));
  pa.scratch_half = pl.scratch_half > 0 ? (pl.scratch_half / (pl.block * pl.ppt)) * block : 0;
  const int DXP = pad_dx(pa.Dx), DYP = pad_dy(pa.Dy);
  const int grid = pa.B;
  ProfScope ps_(PK_PROP, st);
#define PSVO_CASE_P(X, Y)                                                                                           \
  if (DXP == X && DYP == Y) {                                                                                       \
    auto kern = smc_filter_persistent_kernel<X, Y>;                                                                 \
    PSVO_DYN_SMEM(kern, smem);                                                                                      \
    kern<<<grid, block, smem, st>>>(pa, ra, prop_floats);                                                           \
    PSVO_LAUNCH_CHECK();                                                                                            \
    return 0;                                                                                                       \
  }
  PSVO_CASE_P(4, 4) PSVO_CASE_P(4, 8) PSVO_CASE_P(4, 16) PSVO_CASE_P(4, 32)
  PSVO_CASE_P(8, 4) PSVO_CASE_P(8, 8) PSVO_CASE_P(8, 16) PSVO_CASE_P(8, 32)
  PSVO_CASE_P(16, 4) PSVO_CASE_P(16, 8) PSVO_CASE_P(16, 16) PSVO_CASE_P(16, 32)
#undef PSVO_CASE_P
  return fail("unsupported padded dims Dx=%d Dy=%d", pa.Dx, pa.Dy);
}

static bool cluster_env() {
  static const bool off = getenv("PSVO_CLUSTER") != nullptr && getenv("PSVO_CLUSTER")[0] == '0';
  return !off;
}
static size_t cluster_prop_floats(const PropArgs& a) {
  const int DXP = pad_dx(a.Dx), DYP = pad_dy(a.Dy);
  const bool sep_f = !(a.flags & PSVO_F_BOOTSTRAP);
  auto add = [&](int din_pad, int dout_pad, const psvo_mlp& net) {
    if (net.n_hidden == 0) return (size_t)(din_pad * dout_pad + dout_pad);
    return (size_t)net.width[0] * din_pad + round4(net.width[0]) + (size_t)net.width[0] * dout_pad + dout_pad;
  };
  size_t fl = 384 + add(DXP, DXP, a.q1) + (sep_f ? add(DXP, DXP, a.f) : 0) + add(DXP, DYP, a.g);
  return (fl + 3) & ~(size_t)3;
}

// one cluster per sequence pays off for FEW sequences of MANY particles with small networks (the per-step launches are then
// latency); every cluster must be resident at once (one wave), networks with at most one hidden layer
bool cluster_filter_supported(const PropArgs& a, int* csize_out, int* block_out, size_t* smem_out) {
  if (!cluster_env() || !persistent_env()) return false;
  if (a.tc_image != nullptr && tc_propagate_supported(a)) return false;
  const bool sep_f = !(a.flags & PSVO_F_BOOTSTRAP);
  if (a.q1.n_hidden > 1 || a.g.n_hidden > 1 || (sep_f && a.f.n_hidden > 1)) return false;
  if (a.T < 2 || a.N < 512 || a.N > 8192) return false;
  int csize = 8;
  while (csize > 1 && (a.N / csize < 128 || (long long)a.B * csize > num_sms())) csize /= 2;
  if (csize < 2) return false;
  const int chunk = (a.N + csize - 1) / csize;
  {
    // like the persistent kernel: only while one step of one CTA's particles is a small amount of work (with heavy networks
    // the per-step launches spread a sequence over more SMs than a cluster has)
    auto macs = [](const psvo_mlp& net) {
      long long m = 0; int in = net.d_in;
      for (int k = 0; k < net.n_hidden; ++k) { m += (long long)in * net.width[k]; in = net.width[k]; }
      return m + (long long)in * net.d_out;
    };
    const long long instr = 300 + macs(a.q1) + macs(a.g) + (sep_f ? macs(a.f) : 0);
    if ((long long)chunk * instr > 600000) return false;
  }
  int block = chunk < 256 ? ((chunk + 31) & ~31) : 256;
  const size_t total = cluster_prop_floats(a) * sizeof(float) + (size_t)a.N * 12 + (size_t)chunk * 4 + 16;
  if (total > 200 * 1024) return false;
  *csize_out = csize; *block_out = block; *smem_out = total;
  return true;
}

int launch_filter_cluster(PropArgs pa, const ResampleArgs& ra, cudaStream_t st) {
  int csize = 0, block = 0; size_t smem = 0;
  PSVO_REQUIRE(cluster_filter_supported(pa, &csize, &block, &smem), "cluster filter kernel does not support this shape");
  const int prop_floats = (int)cluster_prop_floats(pa);
  pa.scratch_half = 0;
  const int DXP = pad_dx(pa.Dx), DYP = pad_dy(pa.Dy);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pa.B * csize); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  ProfScope ps_(PK_PROP, st);
#define PSVO_CASE_C(X, Y)                                                                                           \
  if (DXP == X && DYP == Y) {                                                                                       \
    auto kern = smc_filter_cluster_kernel<X, Y>;                                                                    \
    PSVO_DYN_SMEM(kern, smem);                                                                                      \
    PSVO_CUDA(cudaLaunchKernelEx(&cfg, kern, pa, ra, prop_floats, csize));                                          \
    PSVO_LAUNCH_CHECK();                                                                                            \
    return 0;                                                                                                       \
  }
  PSVO_CASE_C(4, 4) PSVO_CASE_C(4, 8) PSVO_CASE_C(4, 16) PSVO_CASE_C(4, 32)
  PSVO_CASE_C(8, 4) PSVO_CASE_C(8, 8) PSVO_CASE_C(8, 16) PSVO_CASE_C(8, 32)
  PSVO_CASE_C(16, 4) PSVO_CASE_C(16, 8) PSVO_CASE_C(16, 16) PSVO_CASE_C(16, 32)
#undef PSVO_CASE_C
  return fail("unsupported padded dims Dx=%d Dy=%d", pa.Dx, pa.Dy);
}

int launch_resample(const ResampleArgs& a, cudaStream_t st) {
  const int N = a.N;
  // CTAs per sequence: enough to put ~2 CTAs on every SM, at least 512 children each
  int parts = 1;
  while (parts < 4 && (long long)a.B * parts < 2LL * num_sms() && N / (parts * 2) >= 512) parts *= 2;
  while (parts < 64 && (size_t)N * 12 + (size_t)((N + parts - 1) / parts) * 4 > 220 * 1024) parts *= 2;   // ancestor scratch must fit
  // without the gather (fused into the next propagate launch) one CTA per sequence is best: the CDF is built once and
  // 1024 threads keep four interleaved searches each in flight (measured at config 5: 25 us vs 32 us with 4 parts)
  if (a.skip_gather && (size_t)N * 16 <= 220 * 1024) {
    parts = 1;
    while (parts < 4 && (long long)a.B * parts * 2 < num_sms() && N / (parts * 2) >= 512) parts *= 2;
  }
  {
    // big particle sets with one CTA per sequence: the thread-contiguous kernel (PSVO_RESAMPLE_LEAN=0 keeps the older one)
    // PSVO_RESAMPLE_LEAN=2: test hook -- every launch takes this kernel, one CTA per sequence, whatever N and B
    static const int lean_mode = [] { const char* e = getenv("PSVO_RESAMPLE_LEAN"); return e == nullptr ? 1 : (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)); }();
    const size_t lsm = rg_smem_bytes(N, 1024);
    if (lsm <= 220 * 1024 && (lean_mode == 2 || (lean_mode == 1 && parts == 1 && N >= 2048))) {
      PSVO_DYN_SMEM(smc_resample_lean_kernel, lsm);
      ProfScope ps_(PK_RESAMPLE, st);
      PSVO_CUDA(launch_pdl(smc_resample_lean_kernel, dim3(a.B), dim3(1024), lsm, st, a.pdl != 0, a));
      PSVO_LAUNCH_CHECK();
      return 0;
    }
  }
  const int chunk = (N + parts - 1) / parts;
  int block = N >= 4096 ? (a.skip_gather ? 1024 : 512) : (N >= 1024 ? 256 : (N >= 256 ? 128 : (N > 32 ? 64 : 32)));
  size_t smem = (size_t)N * (sizeof(unsigned long long) + sizeof(float)) + (size_t)chunk * sizeof(int);
  PSVO_REQUIRE(smem <= 220 * 1024, "n_particles=%d needs %zu B of shared memory for the CDF (max 220 KiB)", N, smem);
  // guide table of the inverse-CDF search (K + 1 ints, K = 2^kbits >= N) when it fits and the search is long enough to pay
  int kbits = 0, guide_off = 0;
  while ((1 << kbits) < N) ++kbits;
  // (every CTA of a sequence builds the whole table: with more than two CTAs per sequence it costs what it saves)
  if (N >= 256 && parts <= 2 && kbits <= 24 && smem + ((size_t)(1 << kbits) + 1) * sizeof(int) + 16 <= 220 * 1024) {
    guide_off = (int)((smem + 15) & ~(size_t)15);
    smem = (size_t)guide_off + ((size_t)(1 << kbits) + 1) * sizeof(int);
  }
  PSVO_DYN_SMEM(smc_resample_kernel, smem);
  ProfScope ps_(PK_RESAMPLE, st);
  PSVO_CUDA(launch_pdl(smc_resample_kernel, dim3(a.B, parts), dim3(block), smem, st, a.pdl != 0, a, guide_off, kbits));
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo
