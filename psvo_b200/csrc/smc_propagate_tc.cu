// smc_propagate_tc.cu -- forward-filter "propagate + weight" step with the per-particle MLPs on sm_100a tensor
// cores (tcgen05.mma, TMEM accumulators).  Same contract as smc_propagate_kernel (smc_forward.cu).
//
// Reference: SVO.SMC while_body (src/SMC/SVO.py:123-165) incl. sample_from_2_dist (:182-197,228-232); the two
// per-particle networks are MLP q1 (= f under use_bootstrap) and MLP g (src/transformation/MLP.py:48-54).
//
// Shape of the work (BASELINE config 5: Dx=16, Dy=32, H=256): per particle-step 41 kFLOP of tiny-K GEMMs
//   [128 x 16] x [16 x 256] -> relu -> [128 x 256] x [256 x 16|32].
// One CTA = 4 independent groups of 128 threads.  A group owns one 128-particle tile at a time (TMEM lane == row ==
// thread) and runs a serial pipeline over chunks of 32 hidden units:
//     tensor core : L1(c)   D1[128x32]  = A1 * W1[c]^T          (3 bf16 MMAs: hi*hi + hi*lo + lo*hi, K = 16)
//     threads     : tcgen05.ld D1 row -> + b1 -> ReLU -> bf16 hi/lo -> A2 (shared memory, UMMA K-major layout)
//     tensor core : L2(c)   D2[128xNP] += A2 * W2[c]^T          (2 K-steps x 3 MMAs)   and L1(c+1)
// with ONE mbarrier wait per chunk (tcgen05.commit covers L2(c) and L1(c+1)).  The exposed MMA latency of a group
// is filled by the other three groups' CUDA-core work; the four groups share the staged bf16 weights (80 KB).
// All other arithmetic (fusion, Philox sampling, log densities) is the same FP32 code as the SIMT kernel.
#include <stdlib.h>
#include "mlp_tile.cuh"
#include "smc_kernels.h"
#include "tc_common.cuh"

namespace psvo {

constexpr int kTcGroups = 4;
constexpr int kTcChunk = 32;           // hidden units per pipeline step

struct TcPropSmem {
  uint32_t w1q_hi, w1q_lo, w2q_hi, w2q_lo, w1g_hi, w1g_lo, w2g_hi, w2g_lo;   // bf16 weights (B operands)
  uint32_t b1q, b2q, b1g, b2g, consts;                                        // fp32
  uint32_t a1_hi[kTcGroups], a1_lo[kTcGroups], a2_hi[kTcGroups], a2_lo[kTcGroups];
  uint32_t bars, slot, total;
};

__host__ __device__ inline TcPropSmem tc_prop_layout(int H, int NPG) {
  TcPropSmem s;
  uint32_t o = 0;
#define PSVO_TAKE(bytes) (o += (((bytes) + 127u) & ~127u), o - (((bytes) + 127u) & ~127u))
  s.w1q_hi = PSVO_TAKE(H * 16 * 2); s.w1q_lo = PSVO_TAKE(H * 16 * 2);
  s.w2q_hi = PSVO_TAKE(16 * H * 2); s.w2q_lo = PSVO_TAKE(16 * H * 2);
  s.w1g_hi = PSVO_TAKE(H * 16 * 2); s.w1g_lo = PSVO_TAKE(H * 16 * 2);
  s.w2g_hi = PSVO_TAKE(NPG * H * 2); s.w2g_lo = PSVO_TAKE(NPG * H * 2);
  s.b1q = PSVO_TAKE(H * 4); s.b2q = PSVO_TAKE(16 * 4); s.b1g = PSVO_TAKE(H * 4); s.b2g = PSVO_TAKE(32 * 4);
  s.consts = PSVO_TAKE(128 * 4);
  for (int g = 0; g < kTcGroups; ++g) {
    s.a1_hi[g] = PSVO_TAKE(128 * 16 * 2); s.a1_lo[g] = PSVO_TAKE(128 * 16 * 2);
    s.a2_hi[g] = PSVO_TAKE(128 * kTcChunk * 2); s.a2_lo[g] = PSVO_TAKE(128 * kTcChunk * 2);
  }
  s.bars = PSVO_TAKE(kTcGroups * 8); s.slot = PSVO_TAKE(16);
#undef PSVO_TAKE
  s.total = o;
  return s;
}

struct TcGroup {
  uint32_t a1h, a1l, a2h, a2l;     // shared-memory addresses (u32) of this group's A operands
  unsigned char *pa1h, *pa1l, *pa2h, *pa2l;
  uint64_t* bar;
  uint32_t phase;
  uint32_t tD1, tD2q, tD2g;        // TMEM column bases
  uint32_t lane_base;
  int gt;                          // thread index within the group
  int bar_id;                      // named barrier of the group
};

__device__ __forceinline__ void group_sync(const TcGroup& g) {
  asm volatile("bar.sync %0, 128;\n" :: "r"(g.bar_id) : "memory");
}

// stage this thread's 16-wide row as the A operand of layer 1
__device__ __forceinline__ void tc_stage_row(const TcGroup& g, const float (&x)[16]) {
#pragma unroll
  for (int kg = 0; kg < 2; ++kg) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = x[kg * 8 + i];
    uint4 hi, lo;
    tc::split8(v, hi, lo);
    *reinterpret_cast<uint4*>(g.pa1h + kg * 2048 + g.gt * 16) = hi;
    *reinterpret_cast<uint4*>(g.pa1l + kg * 2048 + g.gt * 16) = lo;
  }
}

__device__ __forceinline__ void tc_issue_l1(const TcGroup& g, uint32_t w1h, uint32_t w1l, int H, int c) {
  const uint32_t idesc = tc::make_idesc_bf16(128, kTcChunk);
  const uint64_t dah = tc::make_desc(g.a1h, 2048, 128), dal = tc::make_desc(g.a1l, 2048, 128);
  const uint64_t dbh = tc::make_desc(w1h + c * (kTcChunk * 16), H * 16, 128);
  const uint64_t dbl = tc::make_desc(w1l + c * (kTcChunk * 16), H * 16, 128);
  tc::mma_bf16(g.tD1, dah, dbh, idesc, false);
  tc::mma_bf16(g.tD1, dah, dbl, idesc, true);
  tc::mma_bf16(g.tD1, dal, dbh, idesc, true);
}

template <int NP>
__device__ __forceinline__ void tc_issue_l2(const TcGroup& g, uint32_t w2h, uint32_t w2l, uint32_t tD2, int c) {
  const uint32_t idesc = tc::make_idesc_bf16(128, NP);
#pragma unroll
  for (int s = 0; s < kTcChunk / 16; ++s) {
    const uint64_t dah = tc::make_desc(g.a2h + s * 4096, 2048, 128), dal = tc::make_desc(g.a2l + s * 4096, 2048, 128);
    const uint32_t boff = (uint32_t)(c * (kTcChunk / 8) + 2 * s) * (NP * 16);
    const uint64_t dbh = tc::make_desc(w2h + boff, NP * 16, 128), dbl = tc::make_desc(w2l + boff, NP * 16, 128);
    tc::mma_bf16(tD2, dah, dbh, idesc, (c > 0) || (s > 0));
    tc::mma_bf16(tD2, dah, dbl, idesc, true);
    tc::mma_bf16(tD2, dal, dbh, idesc, true);
  }
}

// D2[row] = W2^T relu(W1^T x + b1)  for the rows staged in A1 (bias b2 is added by the caller).
// Collective over the group; on return the accumulator tD2 is complete and may be read with tcgen05.ld.
template <int NP>
__device__ __forceinline__ void tc_mlp_group(TcGroup& g, uint32_t w1h, uint32_t w1l, uint32_t w2h, uint32_t w2l,
                                             const float* __restrict__ sb1, int H, uint32_t tD2) {
  tc::fence_proxy_async();
  tc::fence_before_sync();
  group_sync(g);                                   // A1 staged by everyone; previous D2 consumed by everyone
  if (g.gt == 0) {
    tc::fence_after_sync();
    tc_issue_l1(g, w1h, w1l, H, 0);
    tc::mma_commit(g.bar);
  }
  const int nchunk = H / kTcChunk;
  for (int c = 0; c < nchunk; ++c) {
    tc::mbar_wait(g.bar, g.phase); g.phase ^= 1;   // L1(c) done, L2(c-1) done (A2 free)
    tc::fence_after_sync();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[16];
      tc::tmem_ld16(g.tD1 + g.lane_base + half * 16, v);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float h[8];
        const float4 ba = *reinterpret_cast<const float4*>(sb1 + c * kTcChunk + half * 16 + q * 8);
        const float4 bb = *reinterpret_cast<const float4*>(sb1 + c * kTcChunk + half * 16 + q * 8 + 4);
        h[0] = fmaxf(v[q * 8 + 0] + ba.x, 0.f); h[1] = fmaxf(v[q * 8 + 1] + ba.y, 0.f);
        h[2] = fmaxf(v[q * 8 + 2] + ba.z, 0.f); h[3] = fmaxf(v[q * 8 + 3] + ba.w, 0.f);
        h[4] = fmaxf(v[q * 8 + 4] + bb.x, 0.f); h[5] = fmaxf(v[q * 8 + 5] + bb.y, 0.f);
        h[6] = fmaxf(v[q * 8 + 6] + bb.z, 0.f); h[7] = fmaxf(v[q * 8 + 7] + bb.w, 0.f);
        uint4 hi, lo;
        tc::split8(h, hi, lo);
        const int kg = half * 2 + q;
        *reinterpret_cast<uint4*>(g.pa2h + kg * 2048 + g.gt * 16) = hi;
        *reinterpret_cast<uint4*>(g.pa2l + kg * 2048 + g.gt * 16) = lo;
      }
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    group_sync(g);
    if (g.gt == 0) {
      tc::fence_after_sync();
      tc_issue_l2<NP>(g, w2h, w2l, tD2, c);
      if (c + 1 < nchunk) tc_issue_l1(g, w1h, w1l, H, c + 1);
      tc::mma_commit(g.bar);
    }
  }
  tc::mbar_wait(g.bar, g.phase); g.phase ^= 1;     // L2(last) done
  tc::fence_after_sync();
}

template <int NPG>
__global__ void __launch_bounds__(kTcGroups * 128, 1) smc_propagate_tc_kernel(PropArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int Dx = a.Dx, Dy = a.Dy, N = a.N, H = a.g.width[0];
  const bool two_q = a.flags & PSVO_F_TWO_Q, poisson = a.flags & PSVO_F_POISSON;
  const bool t0 = a.t == 0;
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const int tid = threadIdx.x, warp = tid >> 5, grp = warp >> 2;

  float* sb1q = reinterpret_cast<float*>(sm + L.b1q);
  float* sb2q = reinterpret_cast<float*>(sm + L.b2q);
  float* sb1g = reinterpret_cast<float*>(sm + L.b1g);
  float* sb2g = reinterpret_cast<float*>(sm + L.b2g);
  float* C = reinterpret_cast<float*>(sm + L.consts);     // s1[16] | sp[16] | inv1[16] | inv2[16] | sg[32] | ld_p, ld_f, ld_g
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + L.slot);

  // ---- stage weights (bf16 hi/lo, UMMA B layout), biases, constants
  if (!t0) {
    tc::stage_weight_bf16x2(a.q1.W[0], Dx, H, 16, H, sm + L.w1q_hi, sm + L.w1q_lo);
    tc::stage_weight_bf16x2(a.q1.W[1], H, Dx, H, 16, sm + L.w2q_hi, sm + L.w2q_lo);
    for (int i = tid; i < H; i += blockDim.x) sb1q[i] = __ldg(a.q1.b[0] + i);
    for (int i = tid; i < 16; i += blockDim.x) sb2q[i] = i < Dx ? __ldg(a.q1.b[1] + i) : 0.0f;
  }
  tc::stage_weight_bf16x2(a.g.W[0], Dx, H, 16, H, sm + L.w1g_hi, sm + L.w1g_lo);
  tc::stage_weight_bf16x2(a.g.W[1], H, Dy, H, NPG, sm + L.w2g_hi, sm + L.w2g_lo);
  for (int i = tid; i < H; i += blockDim.x) sb1g[i] = __ldg(a.g.b[0] + i);
  for (int i = tid; i < 32; i += blockDim.x) sb2g[i] = i < Dy ? __ldg(a.g.b[1] + i) : 0.0f;
  if (tid < 16) {
    const int d = tid;
    const float s1 = a.sig[(t0 ? SIG_Q0 : SIG_Q1) * 32 + d], s2 = a.sig[SIG_Q2 * 32 + d];
    const float i1 = 1.0f / s1, i2 = 1.0f / s2;
    C[d] = s1; C[16 + d] = two_q ? 1.0f / (i1 + i2) : s1; C[32 + d] = i1; C[48 + d] = i2;
  }
  if (tid >= 32 && tid < 64) C[64 + tid - 32] = a.sig[SIG_G * 32 + tid - 32];
  if (tid < kTcGroups) tc::mbar_init(bars + tid, 1);
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    float lp = 0.f, lf = 0.f, lg = 0.f;
    for (int d = 0; d < Dx; ++d) { lp += logf(C[16 + d]); lf += logf(C[d]); }
    if (!poisson) for (int d = 0; d < Dy; ++d) lg += logf(C[64 + d]);
    C[96] = lp; C[97] = lf; C[98] = lg;
  }
  __syncthreads();
  const float* s1 = C; const float* sp = C + 16; const float* inv1 = C + 32; const float* inv2 = C + 48; const float* sg = C + 64;
  const float ld_p = C[96], ld_f = C[97], ld_g = C[98];
  const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;
  const float neg_logN = -logf((float)N);

  const uint32_t tmem = *slot;
  TcGroup g;
  g.gt = tid & 127;
  g.bar_id = 1 + grp;
  g.bar = bars + grp;
  g.phase = 0;
  g.pa1h = sm + L.a1_hi[grp]; g.pa1l = sm + L.a1_lo[grp]; g.pa2h = sm + L.a2_hi[grp]; g.pa2l = sm + L.a2_lo[grp];
  g.a1h = tc::smem_u32(g.pa1h); g.a1l = tc::smem_u32(g.pa1l); g.a2h = tc::smem_u32(g.pa2h); g.a2l = tc::smem_u32(g.pa2l);
  g.tD1 = tmem + grp * 128; g.tD2q = tmem + grp * 128 + 32; g.tD2g = tmem + grp * 128 + 64;
  g.lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t w1qh = tc::smem_u32(sm + L.w1q_hi), w1ql = tc::smem_u32(sm + L.w1q_lo);
  const uint32_t w2qh = tc::smem_u32(sm + L.w2q_hi), w2ql = tc::smem_u32(sm + L.w2q_lo);
  const uint32_t w1gh = tc::smem_u32(sm + L.w1g_hi), w1gl = tc::smem_u32(sm + L.w1g_lo);
  const uint32_t w2gh = tc::smem_u32(sm + L.w2g_hi), w2gl = tc::smem_u32(sm + L.w2g_lo);

  const int tiles_per_seq = (N + 127) / 128;
  const int work = a.B * tiles_per_seq;
  for (int wi = blockIdx.x * kTcGroups + grp; wi < work; wi += gridDim.x * kTcGroups) {
    const int b = wi / tiles_per_seq;
    const int n = (wi - b * tiles_per_seq) * 128 + g.gt;
    const bool live = n < N;
    const size_t bt = (size_t)b * a.T + a.t;
    const size_t row = bt * N + (live ? n : 0);

    float x[16], mu1[16];
    // ---- d1 mean: MLP q1 on the resampled parent (t >= 1) or the per-sequence q0 mean (t = 0)
    if (!t0) {
      const float* src = a.X_anc + ((bt - 1) * N + (live ? n : 0)) * Dx;
      if (Dx == 16) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 q = reinterpret_cast<const float4*>(src)[v];
          x[4 * v] = q.x; x[4 * v + 1] = q.y; x[4 * v + 2] = q.z; x[4 * v + 3] = q.w;
        }
      } else {
#pragma unroll
        for (int d = 0; d < 16; ++d) x[d] = d < Dx ? src[d] : 0.0f;
      }
      tc_stage_row(g, x);
      tc_mlp_group<16>(g, w1qh, w1ql, w2qh, w2ql, sb1q, H, g.tD2q);
      tc::tmem_ld16(g.tD2q + g.lane_base, mu1);
#pragma unroll
      for (int d = 0; d < 16; ++d) mu1[d] += sb2q[d];
    } else {
#pragma unroll
      for (int d = 0; d < 16; ++d) mu1[d] = d < Dx ? a.mu0[(size_t)b * Dx + d] : 0.0f;
    }

    // ---- proposal (fusion), sample, q / f log densities   (f = d1: use_bootstrap, SVO.py:230)
    float e[16];
    if (a.eps != nullptr) {
      const float* src = a.eps + row * Dx;
#pragma unroll
      for (int d = 0; d < 16; ++d) e[d] = d < Dx ? src[d] : 0.0f;
    } else {
      const uint64_t gidx = ((uint64_t)(a.seq_offset + b) * a.T + a.t) * N + (live ? n : 0);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (4 * c < Dx) normal4(philox4x32_10(a.seed, gidx, c, 0u), &e[4 * c]);
        else { e[4 * c] = e[4 * c + 1] = e[4 * c + 2] = e[4 * c + 3] = 0.f; }
      }
    }
    float qs = 0.f, fsum = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      float m = mu1[d];
      if (two_q) {
        const float m2 = d < Dx ? a.mu2_all[bt * Dx + d] : 0.0f;
        m = sp[d] * (inv1[d] * mu1[d] + inv2[d] * m2);
      }
      const float X = m + sp[d] * e[d];
      const float zq = (X - m) / sp[d];
      const float zf = (X - mu1[d]) / s1[d];
      if (d < Dx) { qs = fmaf(zq, zq, qs); fsum = fmaf(zf, zf, fsum); }
      x[d] = d < Dx ? X : 0.0f;
    }
    const float q_lp = -0.5f * qs - ld_p - cx;
    const float f_lp = -0.5f * fsum - ld_f - cx;
    if (live) {
      float* dst = a.Xs + row * Dx;
      if (Dx == 16) {
#pragma unroll
        for (int v = 0; v < 4; ++v) reinterpret_cast<float4*>(dst)[v] = make_float4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
      } else {
#pragma unroll
        for (int d = 0; d < 16; ++d) if (d < Dx) dst[d] = x[d];
      }
    }

    // ---- emission
    tc_stage_row(g, x);
    tc_mlp_group<NPG>(g, w1gh, w1gl, w2gh, w2gl, sb1g, H, g.tD2g);
    float gsum = 0.f;
#pragma unroll
    for (int c = 0; c < NPG; c += 16) {
      float mug[16];
      tc::tmem_ld16(g.tD2g + g.lane_base + c, mug);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int d = c + i;
        if (d < Dy) {
          const float y = a.obs[bt * Dy + d];
          const float mu = mug[i] + sb2g[d];
          const float z = poisson ? y - (softplus(mu) + 1e-6f) : (y - mu) / sg[d];
          gsum = fmaf(z, z, gsum);
        }
      }
    }
    const float g_lp = -0.5f * gsum - ld_g - cy;
    float lwb = neg_logN;
    if (!t0 && a.logWbar != nullptr && live) lwb = a.logWbar[(size_t)b * N + n];
    if (live) a.log_Ws[bt * N + n] = ((f_lp + g_lp) - q_lp) + lwb;
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

bool tc_propagate_supported(const PropArgs& a) {
  // default flag combination only (f_lp = d1_mvn.log_prob, SVO.py:230); the others stay on the SIMT path
  if (!(a.flags & PSVO_F_BOOTSTRAP) || !(a.flags & PSVO_F_TWO_Q)) return false;
  static const bool disabled = getenv("PSVO_DISABLE_TC") != nullptr && getenv("PSVO_DISABLE_TC")[0] == '1';
  if (disabled) return false;
  if (a.q1.n_hidden != 1 || a.g.n_hidden != 1) return false;
  const int H = a.g.width[0];
  if (a.q1.width[0] != H || H < 32 || H > 256 || (H % 32) != 0) return false;
  if (a.Dx > 16 || a.Dy > 32) return false;
  // the tensor path pays off once the hidden layer dominates and the machine can be filled
  if (H < 128 || (long long)a.B * a.N < 128LL * kTcGroups * 32) return false;
  return true;
}

int launch_propagate_tc(const PropArgs& a, cudaStream_t st) {
  const int H = a.g.width[0];
  const int NPG = a.Dy <= 16 ? 16 : 32;
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const size_t smem = L.total + 128;
  PSVO_REQUIRE(smem <= 227 * 1024, "tensor-core propagate: %zu B of shared memory", smem);
  const int tiles = a.B * ((a.N + 127) / 128);
  int grid = (tiles + kTcGroups - 1) / kTcGroups;
  if (grid > num_sms()) grid = num_sms();
  ProfScope ps_(PK_PROP, st);
  if (NPG == 16) {
    static bool cfg = false;
    if (!cfg) { PSVO_CUDA(cudaFuncSetAttribute(smc_propagate_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
    smc_propagate_tc_kernel<16><<<grid, kTcGroups * 128, smem, st>>>(a);
  } else {
    static bool cfg = false;
    if (!cfg) { PSVO_CUDA(cudaFuncSetAttribute(smc_propagate_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
    smc_propagate_tc_kernel<32><<<grid, kTcGroups * 128, smem, st>>>(a);
  }
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo
