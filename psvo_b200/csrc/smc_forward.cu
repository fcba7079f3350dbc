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
#include <cstdlib>
#include "smc_forward_dev.cuh"

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
