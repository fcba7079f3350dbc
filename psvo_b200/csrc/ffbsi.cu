// ffbsi.cu -- FFBSi-style backward simulation with an M-sample proposal (PSVO) and its "with resampling"
// variant (PSVOwR).
//
// Reference: PSVO.backward_simulation_w_proposal + compute_log_ZSMC (src/SMC/PSVO.py:52-203),
//            PSVOwR.backward_simulation_w_proposal + compute_log_ZSMC (src/SMC/PSVOwR.py:52-198); SURVEY App. A.6.
// Time runs backwards.  Per step:
//   particle_mlp  : mu_f[b][j] = MLP_f(Xs[t-1][j])  (N rows per sequence -- the pairwise term needs only these)
//                   and mu1i[b][i] = MLP_q1_inv(X_{t+1}^{bw}[i]).
//   ffbsi_step    : one thread per candidate (m,i): proposal sample, q / f / g log densities (MLP f and MLP g per
//                   candidate), the PAIRWISE predictive term  log sum_j Wbar_{t-1}[j] f(x~ | x_j)  as an online
//                   log-sum-exp over tiles of mu_f staged in shared memory (the (M,N,N,B,Dx) tensor the reference
//                   materialises with tf.tile, PSVO.py:89,128, never exists), then per trajectory i: normalise
//                   over the M candidates, pick one with the order-independent multinomial rule, write outputs.
//   smc_resample  : (PSVOwR only) second N-way resampling of the picked trajectories (PSVOwR.py:103,145,185).
// A final kernel reduces the stored terms to logZ_b.
#include "mlp_tile.cuh"
#include "smc_kernels.h"

namespace psvo {

int check_dims(const psvo_dims* d);
int check_mlp(const char* name, const psvo_mlp& net, int d_in, int d_out);
SigmaSrc sigma_src(const psvo_dims* d, const psvo_params* p, bool ffbsi);

// out[b][n][:] = MLP(in[b*seq_stride + n*Dx ...]) ; thread per row, weights in shared memory
template <int DIN, int DOUT>
__global__ void __launch_bounds__(256) particle_mlp_kernel(psvo_mlp net, const float* __restrict__ in, long long seq_stride,
                                                           int B, int N, int din, int dout, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  MlpS<DIN, DOUT> s;
  mlp_stage<DIN, DOUT>(net, smem, s);
  griddep_entry();                       // weights staged while the previous kernel drained; `in` is read below
  __syncthreads();
  const long long rows = (long long)B * N;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(r / N), n = (int)(r - (long long)b * N);
    const float* src = in + b * seq_stride + (long long)n * din;
    float x[1][DIN], o[1][DOUT];
#pragma unroll
    for (int d = 0; d < DIN; ++d) x[0][d] = d < din ? src[d] : 0.0f;
    mlp_eval<DIN, DOUT, 1>(s, x, o, nullptr, 0);
#pragma unroll
    for (int d = 0; d < DOUT; ++d) if (d < dout) out[r * dout + d] = o[0][d];
  }
}

static size_t staged_bytes(const psvo_mlp& net, int din_pad, int dout_pad) {
  size_t n;
  if (net.n_hidden == 0) n = (size_t)din_pad * dout_pad + dout_pad;
  else {
    n = (size_t)net.width[0] * din_pad + round4(net.width[0]);
    for (int k = 1; k < net.n_hidden; ++k) n += round4(net.width[k - 1] * net.width[k]) + round4(net.width[k]);
    n += (size_t)net.width[net.n_hidden - 1] * dout_pad + dout_pad;
  }
  return n * sizeof(float);
}

template <int DXP>
static int launch_particle_mlp(const psvo_mlp& net, const float* in, long long seq_stride, int B, int N, int Dx,
                               float* out, cudaStream_t st) {
  PSVO_REQUIRE(net.n_hidden <= 1, "psvo_ffbsi: per-particle networks with more than one hidden layer are not supported yet");
  {
    // many rows through a wide hidden layer: the tcgen05 row MLP (tc_mlp.cu; fp16x3 split, 1e-5 like the other tensor-core paths)
    static const bool tc_off = getenv("PSVO_DISABLE_TC") != nullptr && getenv("PSVO_DISABLE_TC")[0] == '1';
    if (!tc_off && (long long)B * N >= 16384 && net.n_hidden == 1 && net.width[0] >= 128 && net.d_out == Dx && tc_mlp_supported(net))
      return launch_particle_mlp_tc(net, in, seq_stride, B, N, out, st);
  }
  const size_t smem = staged_bytes(net, DXP, DXP);
  auto kern = particle_mlp_kernel<DXP, DXP>;
  PSVO_DYN_SMEM(kern, smem);
  const long long rows = (long long)B * N;
  const int block = rows >= 128 ? 128 : 32;
  long long grid = (rows + block - 1) / block;
  if (grid > 4LL * num_sms()) grid = 4LL * num_sms();
  ProfScope ps_(PK_FFBSI, st);
  PSVO_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(block), smem, st, pdl_enabled(), net, in, seq_stride, B, N, Dx, Dx, out));
  PSVO_LAUNCH_CHECK();
  return 0;
}

// out (B,N,Dx) = MLP(in[b*seq_stride + n*Dx ...]) for a Dx -> Dx network (shared with ffbsi_backward.cu)
int launch_particle_mlp_dx(const psvo_mlp& net, const float* in, long long seq_stride, int B, int N, int Dx, float* out,
                           cudaStream_t st) {
  const int DXP = pad_dx(Dx);
  return DXP == 4 ? launch_particle_mlp<4>(net, in, seq_stride, B, N, Dx, out, st)
                  : (DXP == 8 ? launch_particle_mlp<8>(net, in, seq_stride, B, N, Dx, out, st)
                              : launch_particle_mlp<16>(net, in, seq_stride, B, N, Dx, out, st));
}

struct FfbsiArgs {
  psvo_mlp f, g;
  psvo_mlp q1_inv; float* mu1i_next;   // when mu1i_next != NULL the step also writes MLP q1_inv(picked x) = mu1i of step t-1, in place
                                       // (a particle's own row is read by its candidates before the pick phase writes it)
  const float* sig;
  const float* obs; const float* Xs; const float* log_Ws; const float* lse;
  const float* mu_f; const float* mu1i; const float* m2_all; const float* mu_init_all; const float* mu0i;
  const float* X_next;       // (B,T,N,Dx): bw_Xs (PSVO) or bw_X_anc (PSVOwR); slab t+1 is read
  const float* eps_bw; const float* u_pick; const int* forced_pick;
  float* bw_Xs; int* pick; float* out_f; float* out_g; float* out_a; float* omega_pick; float* omega_all;
  const float* P_ext;        // (B,N,M) pairwise log-sum-exp from the tensor-core kernel, or NULL -> SIMT loop below
  const float* F_ext; const float* G_ext;   // (B,N,M) f_lp / g_lp of every candidate from ffbsi_cand_tc_kernel, or NULL
  const float* Q_ext = nullptr;             // (B,N,M) proposal log-density of every candidate from the same kernel, or NULL
  unsigned long long seed;
  int N, M, B, T, Dx, Dy, t, seq_offset, IL;
  unsigned flags;
  int t_loop = 0;            // 1: ONE launch walks t = T-1 .. 0 itself (mu_f then points at the (T, N, Dx) means of all steps)
};

constexpr int kPairChunk = 128;

template <int DXP, int DYP>
__global__ void __launch_bounds__(256) ffbsi_step_kernel(FfbsiArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int Dx = a.Dx, Dy = a.Dy, N = a.N, M = a.M, T = a.T;
  const bool boot = a.flags & PSVO_F_BOOTSTRAP, two_q = a.flags & PSVO_F_TWO_Q, poisson = a.flags & PSVO_F_POISSON;
  const bool wr = a.flags & PSVO_F_WITH_RESAMPLING;
  // PSVO trajectories never interact in the backward simulation (candidate x~[m,i] depends on trajectory i's own x_{t+1} and on
  // the FILTER's particles only, PSVO.py:116-156), so with the transition means of all steps precomputed a CTA can walk its
  // trajectories through all T steps by itself: the whole pass is ONE launch without any cross-CTA synchronisation
  const bool loop = a.t_loop != 0;

  MlpS<DXP, DXP> fs;
  MlpS<DXP, DYP> gs;
  float* p = smem;
  float* sp = p; p += DXP;  float* inv1 = p; p += DXP;  float* inv2 = p; p += DXP;
  float* sf = p; p += DXP;  float* sinit = p; p += DXP; float* sg = p; p += 32;
  float* lds = p; p += 4;                     // logdet: proposal, f, g, init
  const bool ext = a.F_ext != nullptr;        // candidate densities precomputed on tensor cores: no networks staged here
  if ((loop || a.t != T - 1) && !ext) p = mlp_stage<DXP, DXP>(a.f, p, fs);
  if (!ext) p = mlp_stage<DXP, DYP>(a.g, p, gs);
  MlpS<DXP, DXP> qis;
  if (a.mu1i_next != nullptr && (loop || a.t != 0)) p = mlp_stage<DXP, DXP>(a.q1_inv, p, qis);
  float* sMu = p; p += kPairChunk * DXP;      // mu_f[j] / s_f
  float* sWb = p; p += kPairChunk;            // normalised log weight of j
  float* sCand = p; p += blockDim.x * 5;      // omega | f | g | q | P per candidate
  float* sX = p;                              // [blockDim][DXP] candidate particles

  for (int i = threadIdx.x; i < 32; i += blockDim.x) sg[i] = a.sig[SIG_G * 32 + i];
  griddep_entry();                       // networks (constant over the pass) staged; the previous step's outputs are read below
  const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;
  const float logM = logf((float)M);

  const int IL = a.IL;                                   // trajectories per tile
  const int il = threadIdx.x / M, m = threadIdx.x - il * M;
  const int tiles_per_seq = (N + IL - 1) / IL;
  const int work = a.B * tiles_per_seq;

  for (int wi = blockIdx.x; wi < work; wi += gridDim.x)
  for (int t = loop ? T - 1 : a.t; t >= (loop ? 0 : a.t); --t) {
    const bool last = t == T - 1, first = t == 0;
    const bool fuse_q1i = a.mu1i_next != nullptr && !first;
    const float* mu_f_t = (loop && t >= 1) ? a.mu_f + (size_t)(t - 1) * N * Dx : a.mu_f;
    // per-step constants (the proposal of the last step has its own stddev); the barrier also closes the previous step of
    // this tile: its picks and proposal means (global memory, written by this CTA) are visible to the whole CTA
    __syncthreads();
    if (threadIdx.x < DXP) {
      const int d = threadIdx.x;
      const float s1 = a.sig[(last ? SIG_BQI : SIG_Q1I) * 32 + d], s2 = a.sig[SIG_BQ2 * 32 + d];
      inv1[d] = 1.0f / s1; inv2[d] = 1.0f / s2;
      sp[d] = last ? s1 : 1.0f / (1.0f / s1 + 1.0f / s2);
      sf[d] = a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
      sinit[d] = a.sig[((boot && two_q) ? SIG_Q0 : (boot ? SIG_Q1 : SIG_F)) * 32 + d];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      for (int d = 0; d < Dx; ++d) { l0 += logf(sp[d]); l1 += logf(sf[d]); l3 += logf(sinit[d]); }
      if (!poisson) for (int d = 0; d < Dy; ++d) l2 += logf(sg[d]);
      lds[0] = l0; lds[1] = l1; lds[2] = l2; lds[3] = l3;
    }
    __syncthreads();
    const int b = wi / tiles_per_seq;
    const int i = (wi - b * tiles_per_seq) * IL + il;
    const bool live = il < IL && i < N;
    const size_t bt = (size_t)b * T + t;
    const int ii = live ? i : 0;

    // With every term of omega precomputed on the tensor cores (f, g, q per candidate from ffbsi_cand_tc_kernel, the pairwise
    // term from ffbsi_pair_tc_kernel) this kernel does not need the candidates themselves: only the PICKED one is materialised,
    // after the pick, four dimensions per lane (one Philox block per lane instead of four per candidate)
    const bool lean = ext && a.Q_ext != nullptr && a.P_ext != nullptr && !first && a.mu1i_next == nullptr &&
                      a.forced_pick == nullptr && M >= 4 && M <= 32 && (M & (M - 1)) == 0;
    // ---- proposal sample
    float X[1][DXP], xn[DXP];
    float qs = 0.f;
    if (lean) {
#pragma unroll
      for (int d = 0; d < DXP; ++d) { X[0][d] = 0.0f; xn[d] = 0.0f; }
    } else
#pragma unroll
    for (int c = 0; c < DXP / 4; ++c) {
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (4 * c < Dx) {
        if (a.eps_bw != nullptr) {
          const float* src = a.eps_bw + (((bt * M + (live ? m : 0)) * N + ii) * (size_t)Dx);
#pragma unroll
          for (int k = 0; k < 4; ++k) if (4 * c + k < Dx) e[k] = src[4 * c + k];
        } else {
          const uint64_t gidx = ((((uint64_t)(a.seq_offset + b) * T + t) * M + (live ? m : 0)) * N) + ii;
          normal4(philox4x32_10(a.seed, gidx, c, 2u), e);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = 4 * c + k;
        float mean = 0.f, xnext = 0.f;
        if (d < Dx) {
          if (last) mean = a.mu_init_all[bt * Dx + d];
          else {
            const float m1 = a.mu1i[((size_t)b * N + ii) * Dx + d], m2 = a.m2_all[bt * Dx + d];
            mean = sp[d] * (inv1[d] * m1 + inv2[d] * m2);
            xnext = a.X_next[((bt + 1) * N + ii) * Dx + d];
          }
        }
        const float xv = d < Dx ? mean + sp[d] * e[k] : 0.0f;
        if (d < Dx) { const float z = (xv - mean) / sp[d]; qs = fmaf(z, z, qs); }
        X[0][d] = xv; xn[d] = xnext;
      }
    }
    const float q_lp = lean ? a.Q_ext[((size_t)b * N + ii) * M + (live ? m : 0)] : -0.5f * qs - lds[0] - cx;

    // ---- f(x_{t+1} | x~)  (MLP f on the candidate)
    float f_lp = 0.0f;
    if (ext) {
      f_lp = last ? 0.0f : a.F_ext[((size_t)b * N + ii) * M + (live ? m : 0)];
    } else if (!last) {
      float muf[1][DXP];
      mlp_eval<DXP, DXP, 1>(fs, X, muf, nullptr, 0);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DXP; ++d) if (d < Dx) { const float z = (xn[d] - muf[0][d]) / sf[d]; s = fmaf(z, z, s); }
      f_lp = -0.5f * s - lds[1] - cx;
    }
    // ---- g(y_t | x~)
    float g_lp;
    if (ext) {
      g_lp = a.G_ext[((size_t)b * N + ii) * M + (live ? m : 0)];
    } else {
      float mug[1][DYP];
      mlp_eval<DXP, DYP, 1>(gs, X, mug, nullptr, 0);
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DYP; ++d) if (d < Dy) {
        const float y = a.obs[bt * Dy + d];
        const float z = poisson ? y - (softplus(mug[0][d]) + 1e-6f) : (y - mug[0][d]) / sg[d];
        s = fmaf(z, z, s);
      }
      g_lp = -0.5f * s - lds[2] - cy;
    }
    // ---- predictive term
    float P;
    if (!first && a.P_ext != nullptr) {
      P = a.P_ext[((size_t)b * N + ii) * M + (live ? m : 0)] - lds[1] - cx;
    } else if (!first) {
      float xs[DXP];
#pragma unroll
      for (int d = 0; d < DXP; ++d) xs[d] = d < Dx ? X[0][d] / sf[d] : 0.0f;
      float mx = -INFINITY, sum = 0.0f;
      const float lse_prev = a.lse[bt - 1];
      for (int j0 = 0; j0 < N; j0 += kPairChunk) {
        __syncthreads();
        for (int q = threadIdx.x; q < kPairChunk * DXP; q += blockDim.x) {
          const int jj = q / DXP, d = q - jj * DXP;
          sMu[q] = (j0 + jj < N && d < Dx) ? mu_f_t[((size_t)b * N + j0 + jj) * Dx + d] / sf[d] : 0.0f;
        }
        for (int q = threadIdx.x; q < kPairChunk; q += blockDim.x)
          sWb[q] = j0 + q < N ? a.log_Ws[(bt - 1) * N + j0 + q] - lse_prev : -INFINITY;
        __syncthreads();
        const int nj = min(kPairChunk, N - j0);
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
      P = (mx + logf(sum)) - lds[1] - cx;
    } else {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DXP; ++d) if (d < Dx) { const float z = (X[0][d] - a.mu0i[(size_t)b * Dx + d]) / sinit[d]; s = fmaf(z, z, s); }
      P = -0.5f * s - lds[3] - cx;
    }
    const float omega = last ? (P + g_lp) - q_lp : ((P + f_lp) + g_lp) - q_lp;

    __syncthreads();
    sCand[threadIdx.x * 5 + 0] = omega; sCand[threadIdx.x * 5 + 1] = f_lp; sCand[threadIdx.x * 5 + 2] = g_lp;
    sCand[threadIdx.x * 5 + 3] = q_lp;  sCand[threadIdx.x * 5 + 4] = P;
    if (!lean) {
#pragma unroll
      for (int d = 0; d < DXP; ++d) sX[threadIdx.x * DXP + d] = X[0][d];
    }
    __syncthreads();

    // ---- per trajectory: normalise over the M candidates, pick one  (PSVO.py:96-104,136-144)
    // M a power of two <= 32: a trajectory's candidates are M consecutive lanes of one warp, and the integer part of the pick
    // (2 M deterministic exps when one thread does it alone -- with 30 of 32 lanes idle it was 40 % of this kernel) is one exp
    // per lane, a segmented shuffle scan and a ballot.  Integer sums: the same pick bit for bit.  The floating-point
    // normaliser keeps its sequential order (every lane evaluates it).
    const bool seg_pick = a.forced_pick == nullptr && M <= 32 && (M & (M - 1)) == 0;
    int sel_seg = 0;
    float lse_seg = 0.0f;
    if (seg_pick) {
      const float* cand = sCand + (size_t)il * M * 5;                      // il * M + m = threadIdx.x: always inside the block's array
      float mxo = -INFINITY;
      for (int k = 0; k < M; ++k) mxo = fmaxf(mxo, cand[k * 5]);
      float se = 0.f;
      for (int k = 0; k < M; ++k) se += expf(cand[k * 5] - mxo);
      const float lse_m = mxo + logf(se);
      const float mxn = mxo - lse_m;
      lse_seg = lse_m;
      unsigned long long run = fixed_weight(cand[m * 5] - lse_m, mxn);
      const int sl = threadIdx.x & (M - 1);
      for (int o = 1; o < M; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, run, o, M);
        if (sl >= o) run += up;
      }
      const unsigned long long total = __shfl_sync(0xffffffffu, run, M - 1, M);
      uint32_t u24;
      if (a.u_pick != nullptr) u24 = u24_from_float(a.u_pick[bt * N + ii]);
      else u24 = uniform_bits(a.seed, ((uint64_t)(a.seq_offset + b) * T + t) * N + ii, 3u) >> 8;
      const unsigned long long thr = draw_threshold(total, u24);
      const unsigned vote = __ballot_sync(0xffffffffu, run <= thr);
      const unsigned segbits = (vote >> ((threadIdx.x & 31) & ~(M - 1))) & (M == 32 ? 0xffffffffu : ((1u << M) - 1u));
      sel_seg = __popc(segbits);
      sel_seg = sel_seg < M ? sel_seg : M - 1;
    }
    if (live && m == 0) {
      const float* cand = sCand + (size_t)il * M * 5;
      float mxo = -INFINITY, lse_m = lse_seg;
      if (!seg_pick) {
        for (int k = 0; k < M; ++k) mxo = fmaxf(mxo, cand[k * 5]);
        float se = 0.f;
        for (int k = 0; k < M; ++k) se += expf(cand[k * 5] - mxo);
        lse_m = mxo + logf(se);
      }
      // normalised log omega; its maximum is mxo - lse_m
      const float mxn = mxo - lse_m;
      int sel;
      if (seg_pick) {
        sel = sel_seg;
      } else if (a.forced_pick != nullptr) {
        sel = a.forced_pick[bt * N + i];
        sel = sel < 0 ? 0 : (sel >= M ? M - 1 : sel);
      } else {
        unsigned long long total = 0ull;
        for (int k = 0; k < M; ++k) total += fixed_weight(cand[k * 5] - lse_m, mxn);
        uint32_t u24;
        if (a.u_pick != nullptr) u24 = u24_from_float(a.u_pick[bt * N + i]);
        else u24 = uniform_bits(a.seed, ((uint64_t)(a.seq_offset + b) * T + t) * N + i, 3u) >> 8;
        const unsigned long long thr = draw_threshold(total, u24);
        unsigned long long run = 0ull;
        sel = 0;
        for (int k = 0; k < M; ++k) { run += fixed_weight(cand[k * 5] - lse_m, mxn); if (run <= thr) sel = k + 1; }
        sel = sel < M ? sel : M - 1;
      }
      if (a.omega_all != nullptr)
        for (int k = 0; k < M; ++k) a.omega_all[(bt * N + i) * M + k] = cand[k * 5] - lse_m;
      const float* c = cand + sel * 5;
      const float om = c[0] - lse_m, fl = c[1], gl = c[2], ql = c[3], Pl = c[4];
      const float* xsel = sX + ((size_t)il * M + sel) * DXP;
      if (!lean) for (int d = 0; d < Dx; ++d) a.bw_Xs[(bt * N + i) * Dx + d] = xsel[d];
      a.pick[bt * N + i] = sel;
      if (fuse_q1i) {                                                             // mu1i of step t-1 = MLP q1_inv(bw_X_t) (PSVO.py:120)
        float xin[1][DXP], mo[1][DXP];
#pragma unroll
        for (int d = 0; d < DXP; ++d) xin[0][d] = d < Dx ? xsel[d] : 0.0f;
        mlp_eval<DXP, DXP, 1>(qis, xin, mo, nullptr, 0);
        for (int d = 0; d < Dx; ++d) a.mu1i_next[((size_t)b * N + i) * Dx + d] = mo[0][d];
      }
      if (!wr) {
        a.out_g[bt * N + i] = gl;
        a.out_a[bt * N + i] = (om + ql) + logM;                                   // Omega
        if (!last) a.out_f[(bt + 1) * N + i] = fl;
        if (first) a.out_f[bt * N + i] = Pl;
      } else {
        a.out_a[bt * N + i] = last ? ((((Pl + gl) - ql) - om) - logM) : ((Pl + gl) - ((ql + om) + logM));   // bw_log_W
        a.omega_pick[bt * N + i] = om;
      }
    }
    if (lean && live && 4 * m < Dx) {
      // the picked candidate x~[sel, i] = mean + sp * eps, dimensions [4 m, 4 m + 4) by lane m of the trajectory (same arithmetic
      // as ffbsi_cand_tc_kernel used for its densities)
      const int sel = sel_seg;
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.eps_bw != nullptr) {
        const float* src = a.eps_bw + (((bt * M + sel) * N + i) * (size_t)Dx);
#pragma unroll
        for (int k = 0; k < 4; ++k) if (4 * m + k < Dx) e[k] = src[4 * m + k];
      } else {
        const uint64_t gidx = ((((uint64_t)(a.seq_offset + b) * T + t) * M + sel) * N) + i;
        normal4(philox4x32_10(a.seed, gidx, (uint32_t)m, 2u), e);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = 4 * m + k;
        if (d < Dx) {
          float mean;
          if (last) mean = a.mu_init_all[bt * Dx + d];
          else {
            const float m1 = a.mu1i[((size_t)b * N + i) * Dx + d], m2 = a.m2_all[bt * Dx + d];
            mean = sp[d] * (inv1[d] * m1 + inv2[d] * m2);
          }
          a.bw_Xs[(bt * N + i) * Dx + d] = mean + sp[d] * e[k];
        }
      }
    }
  }
}

// logZ_b: PSVO  logsumexp_n( sum_t (f+g) - sum_t Omega ) - log N ;  PSVOwR  sum_t ( logsumexp_n bw_log_W - log N )
__global__ void __launch_bounds__(256) ffbsi_logz_kernel(const float* __restrict__ f_lp, const float* __restrict__ g_lp,
                                                         const float* __restrict__ oa, int N, int T, int wr,
                                                         float* __restrict__ logZ_b) {
  __shared__ float scr[33];
  const int b = blockIdx.x;
  const float logN = logf((float)N);
  if (!wr) {
    float mx = -INFINITY, sum = 0.0f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float joint = 0.f, prop = 0.f;
      for (int t = 0; t < T; ++t) {
        const size_t k = ((size_t)b * T + t) * N + n;
        joint += f_lp[k] + g_lp[k];
        prop += oa[k];
      }
      const float v = joint - prop;
      if (v > mx) { sum = sum * expf(mx - v) + 1.0f; mx = v; } else sum += expf(v - mx);
    }
    const float gmx = block_max(mx, scr);
    const float part = (mx == -INFINITY) ? 0.0f : sum * expf(mx - gmx);
    const float tot = block_sum(part, scr);
    if (threadIdx.x == 0) logZ_b[b] = gmx + logf(tot) - logN;
  } else {
    float acc = 0.0f;
    for (int t = 0; t < T; ++t) {
      const float* row = oa + ((size_t)b * T + t) * N;
      float mx = -INFINITY;
      for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, row[n]);
      mx = block_max(mx, scr);
      float s = 0.f;
      for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(row[n] - mx);
      s = block_sum(s, scr);
      acc += mx + logf(s) - logN;
    }
    if (threadIdx.x == 0) logZ_b[b] = acc;
  }
}

template <int DXP, int DYP>
static int launch_step_t(const FfbsiArgs& a, int block, int grid, size_t smem, cudaStream_t st) {
  auto kern = ffbsi_step_kernel<DXP, DYP>;
  PSVO_DYN_SMEM(kern, smem);
  ProfScope ps_(PK_FFBSI, st);
  PSVO_CUDA(launch_pdl(kern, dim3(grid), dim3(block), smem, st, pdl_enabled(), a));
  PSVO_LAUNCH_CHECK();
  return 0;
}

struct FfWs { float *mu_f_all; float *sig, *mu_f, *mu1i, *m2_all, *mu_init_all, *mu0i, *omega_pick, *candP, *candF, *candG, *candQ; unsigned char *bimg, *tc_image, *tc_image2; int bimg_steps; size_t bytes; };
static FfWs carve_ff(const psvo_dims* d, const psvo_params* p, char* base) {
  FfWs w; size_t off = 0;
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(base + off); off += align_up(floats * sizeof(float), 256); return r; };
  const size_t BT = (size_t)d->B * d->T, BN = (size_t)d->B * d->N;
  w.sig = take(SIG_SLOTS * 32);
  w.mu_f = take(BN * d->Dx); w.mu1i = take(BN * d->Dx);
  // single-sequence batches (the reference's own operating point, batch_size = 1) are launch-latency bound: the transition means
  // of ALL time steps (MLP f on the filter's particles, known up front) are then computed by one launch instead of T
  w.mu_f_all = (d->B == 1 && (size_t)d->T * d->N * d->Dx <= (size_t)1 << 26) ? take((size_t)d->T * d->N * d->Dx) : nullptr;
  w.m2_all = take(BT * d->Dx); w.mu_init_all = take(BT * d->Dx); w.mu0i = take((size_t)d->B * d->Dx);
  w.omega_pick = take(BT * d->N);
  w.candP = nullptr; w.bimg = nullptr; w.bimg_steps = 1;
  if (ffbsi_pair_tc_supported(d->N, d->M, d->B, d->Dx)) {
    w.candP = take(BN * d->M);
    // batch 1: one B image per time step (they depend on forward outputs only), prepared by ONE launch before the loop
    const size_t ib = ffbsi_pair_tc_image_bytes(d->B, d->N);
    w.bimg_steps = (w.mu_f_all != nullptr && d->T >= 2 && d->T <= 65536 && (size_t)d->T * ib <= (size_t)64 << 20) ? d->T : 1;
    w.bimg = reinterpret_cast<unsigned char*>(take((size_t)w.bimg_steps * ib / sizeof(float)));
  }
  w.candF = w.candG = w.candQ = nullptr; w.tc_image = w.tc_image2 = nullptr;
  if (p != nullptr) {
    const psvo_mlp& fnet = (d->flags & PSVO_F_BOOTSTRAP) ? p->q1.net : p->f.net;
    if (ffbsi_cand_tc_supported(fnet, p->g.net, d->Dx, d->Dy, (long long)BN * d->M)) {
      w.candF = take(BN * d->M); w.candG = take(BN * d->M); w.candQ = take(BN * d->M);
      w.tc_image = reinterpret_cast<unsigned char*>(take(32 * 1024));
      w.tc_image2 = reinterpret_cast<unsigned char*>(take(32 * 1024));      // q1_inv in the "q" slot: the row MLPs of the time loop
    }
  }
  w.bytes = off;
  return w;
}

}  // namespace psvo

using namespace psvo;

extern "C" {

PSVO_API size_t psvo_ffbsi_workspace_bytes(const psvo_dims* d, const psvo_params* p) {
  (void)p;
  if (d == nullptr) return 0;
  return carve_ff(d, p, nullptr).bytes;
}

PSVO_API int psvo_ffbsi_forward(const psvo_dims* d, const psvo_params* p, const float* obs, const float* bs_feat,
                                const float* pre_x0, const float* Xs, const float* log_Ws, const float* lse,
                                const float* eps_bw, const float* u_pick, const float* u_res, const int32_t* forced_pick,
                                const int32_t* forced_anc2, uint64_t seed, float* bw_Xs, int32_t* pick, float* f_lp,
                                float* g_lp, float* Omega_or_bw_log_W, float* bw_X_anc, int32_t* anc2, float* omega_all,
                                float* logZ_b, void* ws, size_t ws_bytes, psvo_stream_t stream) {
  if (int e = check_dims(d)) return e;
  PSVO_REQUIRE(p != nullptr, "params is null");
  const bool boot = d->flags & PSVO_F_BOOTSTRAP, two_q = d->flags & PSVO_F_TWO_Q, wr = d->flags & PSVO_F_WITH_RESAMPLING;
  PSVO_REQUIRE(d->M >= 2 && d->M <= 256, "n_particles_for_BSim_proposal=%d must be in [2,256] (the reference asserts M >= 2, src/SMC/SVO.py:259)", d->M);
  PSVO_REQUIRE(d->T >= 2, "backward simulation needs time >= 2");
  PSVO_REQUIRE(d->Dfeat >= 1, "Dfeat must be positive");
  const psvo_dist& fd = boot ? p->q1 : p->f;
  const psvo_dist& init = (boot && two_q) ? p->q0 : fd;
  if (int e = check_mlp("f", fd.net, d->Dx, d->Dx)) return e;
  if (int e = check_mlp("g", p->g.net, d->Dx, d->Dy)) return e;
  if (int e = check_mlp("BSim_q_init", p->bsim_q_init.net, d->Dfeat, d->Dx)) return e;
  if (int e = check_mlp("q1_inv", p->q1_inv.net, d->Dx, d->Dx)) return e;
  if (int e = check_mlp("BSim_q2", p->bsim_q2.net, d->Dfeat, d->Dx)) return e;
  if (int e = check_mlp("init dist", init.net, d->Dx0, d->Dx)) return e;
  PSVO_REQUIRE(fd.net.n_hidden <= 1 && p->g.net.n_hidden <= 1 && p->q1_inv.net.n_hidden <= 1,
               "psvo_ffbsi_forward: per-particle networks with more than one hidden layer are not supported yet");
  PSVO_REQUIRE(obs && bs_feat && pre_x0 && Xs && log_Ws && lse && bw_Xs && pick && Omega_or_bw_log_W && logZ_b && ws,
               "psvo_ffbsi_forward: null buffer");
  PSVO_REQUIRE(wr ? (bw_X_anc && anc2) : (f_lp && g_lp), "psvo_ffbsi_forward: missing %s outputs", wr ? "PSVOwR" : "PSVO");
  const FfWs w = carve_ff(d, p, static_cast<char*>(ws));
  PSVO_REQUIRE(ws_bytes >= w.bytes, "workspace too small: %zu < %zu", ws_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int B = d->B, T = d->T, N = d->N, M = d->M, Dx = d->Dx;
  const size_t BT = (size_t)B * T;

  if (int e = launch_sigma_block(sigma_src(d, p, true), w.sig, st)) return e;
  if (int e = launch_mlp_rows_forward(p->bsim_q2.net, bs_feat, (int64_t)BT, w.m2_all, st)) return e;
  if (int e = launch_mlp_rows_forward(p->bsim_q_init.net, bs_feat, (int64_t)BT, w.mu_init_all, st)) return e;
  if (int e = launch_mlp_rows_forward(init.net, pre_x0, B, w.mu0i, st)) return e;

  FfbsiArgs a;
  a.f = fd.net; a.g = p->g.net; a.sig = w.sig; a.obs = obs; a.Xs = Xs; a.log_Ws = log_Ws; a.lse = lse;
  a.mu_f = w.mu_f; a.mu1i = w.mu1i; a.m2_all = w.m2_all; a.mu_init_all = w.mu_init_all; a.mu0i = w.mu0i;
  a.X_next = wr ? bw_X_anc : bw_Xs;
  a.eps_bw = eps_bw; a.u_pick = u_pick; a.forced_pick = forced_pick;
  a.bw_Xs = bw_Xs; a.pick = pick; a.out_f = f_lp; a.out_g = g_lp; a.out_a = Omega_or_bw_log_W;
  a.omega_pick = w.omega_pick; a.omega_all = omega_all; a.seed = seed;
  a.P_ext = nullptr;
  a.q1_inv = p->q1_inv.net; a.mu1i_next = nullptr;
  const bool pair_tc = w.candP != nullptr;
  PairTcArgs pt;
  pt.sig = w.sig; pt.mu_f = w.mu_f; pt.mu1i = w.mu1i; pt.m2_all = w.m2_all; pt.mu_init_all = w.mu_init_all;
  pt.log_Ws = log_Ws; pt.lse = lse; pt.eps_bw = eps_bw; pt.bimg = w.bimg; pt.candP = w.candP; pt.seed = seed;
  pt.N = N; pt.M = M; pt.B = B; pt.T = T; pt.Dx = Dx; pt.seq_offset = d->seq_offset; pt.flags = d->flags;
  a.N = N; a.M = M; a.B = B; a.T = T; a.Dx = Dx; a.Dy = d->Dy; a.seq_offset = d->seq_offset; a.flags = d->flags;

  // candidate MLPs f and g on tensor cores (one fp16 weight image per pass, f in the "q" slot)
  a.F_ext = nullptr; a.G_ext = nullptr;
  const bool cand_tc = w.candF != nullptr;
  bool rows_tc = false;      // the two row MLPs of a step (transition means, inverse-proposal means) on the tcgen05 pipeline too
  CandTcArgs ca;
  if (cand_tc) {
    PropArgs ia;
    ia.q1 = fd.net; ia.f = fd.net; ia.g = p->g.net; ia.Dx = Dx; ia.Dy = d->Dy;
    if (int e = launch_tc_stage_image(ia, w.tc_image, st)) return e;
    if (tc_pipeline_supported(p->q1_inv.net, p->g.net, Dx, d->Dy)) {
      PropArgs ib = ia;
      ib.q1 = p->q1_inv.net; ib.f = p->q1_inv.net;
      if (int e = launch_tc_stage_image(ib, w.tc_image2, st)) return e;
      rows_tc = true;
    }
    ca.tc_image = w.tc_image; ca.sig = w.sig; ca.obs = obs; ca.mu1i = w.mu1i; ca.m2_all = w.m2_all; ca.mu_init_all = w.mu_init_all;
    ca.X_next = a.X_next; ca.eps_bw = eps_bw; ca.candF = w.candF; ca.candG = w.candG; ca.candQ = w.candQ; ca.seed = seed;
    ca.N = N; ca.M = M; ca.B = B; ca.T = T; ca.Dx = Dx; ca.Dy = d->Dy; ca.H = p->g.net.width[0]; ca.seq_offset = d->seq_offset;
    ca.flags = d->flags;
    a.F_ext = w.candF; a.G_ext = w.candG; a.Q_ext = w.candQ;
  }

  const int DXP = pad_dx(Dx), DYP = pad_dy(d->Dy);
  int IL = 256 / M; if (IL < 1) IL = 1;
  if (IL > N) IL = N;
  // few tiles (small batches): one warp per CTA on as many SMs as there are tiles beats eight warps sharing one SM and its
  // barriers -- the step is a dependent chain per warp (config 2: 12.4 -> 12.1 ms per pass on the same box).  Per-trajectory
  // arithmetic does not depend on the tile shape.  PSVO_FFBSI_IL caps the trajectories per tile by hand.
  while (IL * M > 32 && IL % 2 == 0 && (int64_t)B * ((N + IL - 1) / IL) < num_sms()) IL /= 2;
  {
    static const int il_env = getenv("PSVO_FFBSI_IL") ? atoi(getenv("PSVO_FFBSI_IL")) : 0;
    if (il_env > 0 && il_env <= N && il_env * M <= 256) IL = il_env;
  }
  int block = ((IL * M + 31) / 32) * 32;
  a.IL = IL;
  // PSVO on the FP32 step kernel: MLP q1_inv of the picked particle (the next step's proposal mean) is evaluated by the step
  // kernel itself instead of a launch of its own per time step
  const bool fuse_q1i = !wr && !cand_tc;
  if (fuse_q1i) a.mu1i_next = w.mu1i;
  const size_t smem = (5 * DXP + 32 + 4 + kPairChunk * DXP + kPairChunk + (size_t)block * 5 + (size_t)block * DXP) * sizeof(float) +
                      (cand_tc ? 0 : staged_bytes(a.f, DXP, DXP) + staged_bytes(a.g, DXP, DYP)) +
                      (fuse_q1i ? staged_bytes(a.q1_inv, DXP, DXP) : 0);
  PSVO_REQUIRE(smem <= 227 * 1024, "psvo_ffbsi_forward: networks too large for shared memory (%zu B)", smem);
  const int64_t work = (int64_t)B * ((N + IL - 1) / IL);
  const int64_t cap = (int64_t)num_sms() * (smem > 100 * 1024 ? 1 : 2);
  const int grid = (int)(work < cap ? work : cap);

  ResampleArgs ra;
  ra.log_Ws = w.omega_pick; ra.Xs = bw_Xs; ra.u = u_res; ra.forced_anc = forced_anc2; ra.X_anc = bw_X_anc; ra.anc = anc2;
  ra.lse = nullptr; ra.logZ_b = nullptr; ra.logWbar = nullptr; ra.seed = seed;
  ra.N = N; ra.B = B; ra.T = T; ra.Dx = Dx; ra.seq_offset = d->seq_offset; ra.flags = PSVO_F_RESAMPLE; ra.stream_id = 4u;

  if (w.mu_f_all != nullptr) {     // B == 1: row (t, n) of Xs -> row (t, n) of mu_f_all ("sequences" = time steps)
    const int rc0 = DXP == 4 ? launch_particle_mlp<4>(fd.net, Xs, (long long)N * Dx, T, N, Dx, w.mu_f_all, st)
                             : (DXP == 8 ? launch_particle_mlp<8>(fd.net, Xs, (long long)N * Dx, T, N, Dx, w.mu_f_all, st)
                                         : launch_particle_mlp<16>(fd.net, Xs, (long long)N * Dx, T, N, Dx, w.mu_f_all, st));
    if (rc0) return rc0;
  }
  // PSVO on the FP32 kernels with the transition means of all steps at hand (batch 1): the whole backward simulation is ONE
  // launch, every CTA walking its trajectories through all T steps (PSVO_FFBSI_LOOP=0 keeps one launch per step)
  static const bool loop_off = getenv("PSVO_FFBSI_LOOP") != nullptr && getenv("PSVO_FFBSI_LOOP")[0] == '0';
  const bool whole_pass = !loop_off && fuse_q1i && !pair_tc && !cand_tc && w.mu_f_all != nullptr && T >= 2;
  if (whole_pass) {
    a.t = T - 1; a.t_loop = 1; a.mu_f = w.mu_f_all; a.P_ext = nullptr;
    int rc = -1;
#define PSVO_CASE_FL(X, Y) if (DXP == X && DYP == Y) rc = launch_step_t<X, Y>(a, block, grid, smem, st);
    PSVO_CASE_FL(4, 4) PSVO_CASE_FL(4, 8) PSVO_CASE_FL(4, 16) PSVO_CASE_FL(4, 32)
    PSVO_CASE_FL(8, 4) PSVO_CASE_FL(8, 8) PSVO_CASE_FL(8, 16) PSVO_CASE_FL(8, 32)
    PSVO_CASE_FL(16, 4) PSVO_CASE_FL(16, 8) PSVO_CASE_FL(16, 16) PSVO_CASE_FL(16, 32)
#undef PSVO_CASE_FL
    if (rc) return rc;
  }
  if (pair_tc && w.bimg_steps > 1) {
    PairTcArgs pa = pt;
    pa.bimg = w.bimg; pa.mu_f = w.mu_f_all;
    if (int e = launch_ffbsi_pair_prep_all(pa, st)) return e;
  }
  for (int t = whole_pass ? -1 : T - 1; t >= 0; --t) {
    a.t = t;
    int rc = 0;
    if (t >= 1 && w.mu_f_all != nullptr) {
      a.mu_f = pt.mu_f = w.mu_f_all + (size_t)(t - 1) * N * Dx;
    } else if (t >= 1 && rows_tc) {
      Fwd2TcArgs fa;
      fa.tc_image = w.tc_image; fa.xq = Xs + (size_t)(t - 1) * N * Dx; fa.xg = nullptr; fa.x_seq_stride = (long long)T * N * Dx;
      fa.muq = w.mu_f; fa.mug = nullptr; fa.B = B; fa.N = N; fa.Dx = Dx; fa.Dy = d->Dy; fa.H = p->g.net.width[0]; fa.prof_kind = PK_FFBSI;
      if ((rc = launch_tc_fwd2(fa, st))) return rc;
    } else if (t >= 1) {      // mu_f = MLP_f(Xs[t-1])
      const float* in = Xs + (size_t)(t - 1) * N * Dx;
      rc = DXP == 4 ? launch_particle_mlp<4>(fd.net, in, (long long)T * N * Dx, B, N, Dx, w.mu_f, st)
                    : (DXP == 8 ? launch_particle_mlp<8>(fd.net, in, (long long)T * N * Dx, B, N, Dx, w.mu_f, st)
                                : launch_particle_mlp<16>(fd.net, in, (long long)T * N * Dx, B, N, Dx, w.mu_f, st));
      if (rc) return rc;
    }
    if (t <= T - 2 && !fuse_q1i && rows_tc) {
      Fwd2TcArgs fa;
      fa.tc_image = w.tc_image2; fa.xq = a.X_next + (size_t)(t + 1) * N * Dx; fa.xg = nullptr; fa.x_seq_stride = (long long)T * N * Dx;
      fa.muq = w.mu1i; fa.mug = nullptr; fa.B = B; fa.N = N; fa.Dx = Dx; fa.Dy = d->Dy; fa.H = p->g.net.width[0]; fa.prof_kind = PK_FFBSI;
      if ((rc = launch_tc_fwd2(fa, st))) return rc;
    } else if (t <= T - 2 && !fuse_q1i) {  // mu1i = MLP_q1_inv(X_{t+1}^{bw})
      const float* in = a.X_next + (size_t)(t + 1) * N * Dx;
      rc = DXP == 4 ? launch_particle_mlp<4>(p->q1_inv.net, in, (long long)T * N * Dx, B, N, Dx, w.mu1i, st)
                    : (DXP == 8 ? launch_particle_mlp<8>(p->q1_inv.net, in, (long long)T * N * Dx, B, N, Dx, w.mu1i, st)
                                : launch_particle_mlp<16>(p->q1_inv.net, in, (long long)T * N * Dx, B, N, Dx, w.mu1i, st));
      if (rc) return rc;
    }
    if (cand_tc) {
      ca.t = t;
      if ((rc = launch_ffbsi_cand_tc(ca, st))) return rc;
    }
    a.P_ext = nullptr;
    if (pair_tc && t >= 1) {
      pt.t = t;
      if (w.bimg_steps > 1) { pt.bimg = w.bimg + (size_t)t * ffbsi_pair_tc_image_bytes(B, N); pt.prep_done = 1; }
      if ((rc = launch_ffbsi_pair_tc(pt, st))) return rc;
      a.P_ext = w.candP;
    }
    rc = -1;
#define PSVO_CASE_F(X, Y) if (DXP == X && DYP == Y) rc = launch_step_t<X, Y>(a, block, grid, smem, st);
    PSVO_CASE_F(4, 4) PSVO_CASE_F(4, 8) PSVO_CASE_F(4, 16) PSVO_CASE_F(4, 32)
    PSVO_CASE_F(8, 4) PSVO_CASE_F(8, 8) PSVO_CASE_F(8, 16) PSVO_CASE_F(8, 32)
    PSVO_CASE_F(16, 4) PSVO_CASE_F(16, 8) PSVO_CASE_F(16, 16) PSVO_CASE_F(16, 32)
#undef PSVO_CASE_F
    if (rc) return rc;
    if (wr) {
      ra.t = t;
      if (int e = launch_resample(ra, st)) return e;
    }
  }
  {
    ProfScope ps_(PK_FFBSI, st);
    ffbsi_logz_kernel<<<B, 256, 0, st>>>(f_lp, g_lp, Omega_or_bw_log_W, N, T, wr ? 1 : 0, logZ_b);
    PSVO_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
