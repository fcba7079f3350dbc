// smc_persistent.cu -- persistent whole-filter kernel: ONE launch per forward pass, one CTA per sequence looping over time
// (device code: smc_forward_dev.cuh).  Reference: the tf.while_loop of SVO.SMC (src/SMC/SVO.py:123-165).
#include <stdlib.h>
#include <cstdlib>
#include "smc_forward_dev.cuh"

namespace psvo {

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

// MAXB: the block-size bound the instantiation is compiled for -- 1024 threads leave 64 registers per thread (spills in the
// propagate tile), so launches of at most 512 / 256 threads get their own instantiations with 128 / 255
template <int DXP, int DYP, int MAXB = persist_max_block(DXP, DYP)>
__global__ void __launch_bounds__(MAXB, 1) smc_filter_persistent_kernel(PropArgs pa, ResampleArgs ra, int prop_floats) {
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

static size_t persist_resample_bytes(int N, int block) {
  if (N < kPersistLeanMinN) return (size_t)N * 16;
  return ((rg_smem_bytes(N, block) + 15) & ~(size_t)15) + (size_t)N * sizeof(int);
}

bool persistent_env() {
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
  // the register-rich instantiations hold one CTA per SM: only while every sequence still gets an SM of its own (config 3, 256
  // sequences, needs two co-resident CTAs per SM: 2.14 vs 1.77 ms per pass with them)
  const bool roomy = pa.B <= num_sms();
  ProfScope ps_(PK_PROP, st);
#define PSVO_CASE_PB(X, Y, MB)                                                                                      \
  {                                                                                                                 \
    auto kern = smc_filter_persistent_kernel<X, Y, MB>;                                                             \
    PSVO_DYN_SMEM(kern, smem);                                                                                      \
    kern<<<grid, block, smem, st>>>(pa, ra, prop_floats);                                                           \
    PSVO_LAUNCH_CHECK();                                                                                            \
    return 0;                                                                                                       \
  }
#define PSVO_CASE_P(X, Y)                                                                                           \
  if (DXP == X && DYP == Y) {                                                                                       \
    if (roomy && persist_max_block(X, Y) > 256 && block <= 256) PSVO_CASE_PB(X, Y, 256)                             \
    if (roomy && persist_max_block(X, Y) > 512 && block <= 512) PSVO_CASE_PB(X, Y, 512)                             \
    PSVO_CASE_PB(X, Y, persist_max_block(X, Y))                                                                     \
  }
  PSVO_CASE_P(4, 4) PSVO_CASE_P(4, 8) PSVO_CASE_P(4, 16) PSVO_CASE_P(4, 32)
  PSVO_CASE_P(8, 4) PSVO_CASE_P(8, 8) PSVO_CASE_P(8, 16) PSVO_CASE_P(8, 32)
  PSVO_CASE_P(16, 4) PSVO_CASE_P(16, 8) PSVO_CASE_P(16, 16) PSVO_CASE_P(16, 32)
#undef PSVO_CASE_P
#undef PSVO_CASE_PB
  return fail("unsupported padded dims Dx=%d Dy=%d", pa.Dx, pa.Dy);
}

}  // namespace psvo
