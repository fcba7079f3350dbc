// smc_cluster.cu -- thread-block-cluster whole-filter kernel: ONE launch per forward pass, one cluster per sequence
// (device code: smc_forward_dev.cuh).  Reference: the tf.while_loop of SVO.SMC (src/SMC/SVO.py:123-165).
#include <stdlib.h>
#include <cstdlib>
#include "smc_forward_dev.cuh"

namespace psvo {

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
  if (const char* e = getenv("PSVO_CLUSTER_SIZE")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) csize = v; }   // experiment hook
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

}  // namespace psvo
