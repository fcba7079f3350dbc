// ffbsi_pair_tc.cu -- the PAIRWISE predictive term of the FFBSi backward simulation on sm_100a tensor cores.
//
// Reference: PSVO.backward_simulation_w_proposal, src/SMC/PSVO.py:89-94,128-133 (and PSVOwR.py:88-93,127-133):
//     f_tm1_log_prob = f.log_prob(Xs[t-1], tile(bw_X_t))          (M, N, N, B)   <- materialised by tf.tile
//     log_W_t        = reduce_logsumexp(f_tm1_log_prob + log_W_tm1, axis=2)
// i.e. for every candidate row r = (i, m) of a sequence:  P[r] = LSE_j( log N(x~_r; mu_j, s_f) + log Wbar_{t-1}[j] ).
//
// As a GEMM: with xs = x~/s_f, mus = mu_j/s_f,
//     log N(...) + log Wbar_j = const + ( xs . mus + c_j ) - |xs|^2 / 2 ,     c_j = log Wbar_j - |mus|^2 / 2
// rows = candidates (M*N per sequence), columns = filter particles j (N), K = Dx <= 16.  The c_j term rides in the
// GEMM as two extra K entries (A = 1, B = c_j split hi/lo), the row term is added after the log-sum-exp.
//
// Arithmetic: fp16x3 split products (tc_common.cuh) with both operands pre-scaled by 16, fp32 accumulation in TMEM:
//     D = 256 * (xs . mus + c_j)  =  hi*hi + [256 * c_hi + 256 * c_lo] + hi*lo + lo*hi          (4 MMAs, N = 128)
//
// CTA (persistent, one per SM) = 2 compute groups of 128 threads (TMEM lane == candidate row == thread), one
// MMA-issuing warp, one TMA producer warp:
//     producer : streams the per-sequence B image (128-column tiles, 12 KB each, prepared by ffbsi_pair_prep_kernel in
//                the UMMA operand layout) through a 4-stage ring with cp.async.bulk + mbarrier transaction counts;
//     MMA warp : per column tile and group, 4 tcgen05.mma into a double-buffered 128-column accumulator; tcgen05.commit
//                releases the B stage and publishes the accumulator;
//     compute  : build the A operand of their row tile once (recompute the candidate from eps_bw / Philox), then per
//                column tile tcgen05.ld their row and fold it into an online log-sum-exp (1 ex2 per pair -- the MUFU
//                pipe, not the tensor pipe, bounds this kernel: 16 ex2/clk/SM).
// The (M,N,N,B,Dx) tensor never exists; the tile never leaves the SM; HBM traffic is the 4-byte P per candidate.
#include <stdlib.h>
#include "smc_kernels.h"
#include "tc_common.cuh"

namespace psvo {

constexpr int kPtGroups = 4;             // compute groups (128 candidate rows each) per CTA: 16 warps keep the MUFU pipe fed
constexpr int kPtCols = 64;              // columns (filter particles j) per tile = accumulator width; 4 groups x 2 buffers x 64 = 512 TMEM columns
constexpr int kPtStages = 6;
constexpr int kPtKg = kPtCols * 16;      // bytes of one K-group (8 k) of a B tile
constexpr int kPtTileBytes = 6 * kPtKg;  // hi+ext K=32 (4 K-groups) | lo K=16 (2 K-groups)
constexpr int kPtABytes = 12288;         // A operand of one group: 128 rows, same K-group structure (2 KB per K-group)
constexpr int kPtThreads = kPtGroups * 128 + 96;   // + 2 MMA warps (groups 0-1, 2-3) + 1 TMA producer warp
constexpr float kPtScale = 16.0f;        // operand scale (exact); the accumulator holds 256 * (dot + c_j)
constexpr float kPtSentinel = -60000.0f; // c_j of padded columns / zero-weight particles (finite in fp16)

// ---- B image: per (sequence, 128-column tile) the UMMA B operand [N = 128 rows j][K] ----------------------------
__global__ void __launch_bounds__(128) ffbsi_pair_prep_kernel(PairTcArgs a) {
  griddep_entry();
  const int b = blockIdx.y;
  const int N = a.N, Dx = a.Dx, NT = (N + kPtCols - 1) / kPtCols;
  const bool boot = a.flags & PSVO_F_BOOTSTRAP;
  const int j = blockIdx.x * 128 + threadIdx.x;
  const int ct = j / kPtCols, r = j - ct * kPtCols;
  if (ct >= NT) return;
  const bool live = j < N;
  const int t = a.prep_all ? 1 + (int)blockIdx.z : a.t;                       // prep_all: grid.z walks the steps 1..T-1 (batch 1)
  const size_t btm = (size_t)b * a.T + (t - 1);
  const float* mu_f_t = a.prep_all ? a.mu_f + (size_t)(t - 1) * N * Dx : a.mu_f;
  float mus[16];
  float nrm = 0.0f;
#pragma unroll
  for (int d = 0; d < 16; ++d) {
    float v = 0.0f;
    if (live && d < Dx) v = mu_f_t[((size_t)b * N + j) * Dx + d] / a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
    mus[d] = v;
    nrm = fmaf(v, v, nrm);
  }
  float cj = kPtSentinel;
  if (live) cj = fmaxf((a.log_Ws[btm * N + j] - a.lse[btm]) - 0.5f * nrm, kPtSentinel);
  unsigned char* img = a.bimg + (a.prep_all ? (size_t)t * a.B * NT * kPtTileBytes : 0) + ((size_t)b * NT + ct) * kPtTileBytes;
#pragma unroll
  for (int kg = 0; kg < 2; ++kg) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = mus[kg * 8 + i] * kPtScale;
    uint4 hi, lo;
    tc::split8_f16(v, hi, lo);
    *reinterpret_cast<uint4*>(img + kg * kPtKg + r * 16) = hi;
    *reinterpret_cast<uint4*>(img + 4 * kPtKg + kg * kPtKg + r * 16) = lo;
  }
  const __half ch = __float2half_rn(cj);
  const __half cl = __float2half_rn(cj - __half2float(ch));
  const __half2 c2 = __halves2half2(ch, cl);
  *reinterpret_cast<uint4*>(img + 2 * kPtKg + r * 16) = make_uint4(*reinterpret_cast<const uint32_t*>(&c2), 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(img + 3 * kPtKg + r * 16) = make_uint4(0u, 0u, 0u, 0u);
}

// one MUFU instruction: 2^x, flush-to-zero (arguments are <= 0 here; relative error ~2^-22)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe (x <= 0): round-to-nearest split x = n + f, |f| <= 1/2, degree-6 Taylor of 2^f (relative error
// 1.2e-7), exponent patched with integer arithmetic.  One pair in four goes this way so that the MUFU pipe -- the bound of
// this kernel -- and the FMA pipe work side by side (the trick of FlashAttention-4's software exp).
__device__ __forceinline__ float poly_ex2(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = 1.5403530e-4f;
  p = fmaf(p, f, 1.3333558e-3f);
  p = fmaf(p, f, 9.6181291e-3f);
  p = fmaf(p, f, 5.5504109e-2f);
  p = fmaf(p, f, 2.4022651e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

struct PtSmem { uint32_t stage, aop, bars, slot, consts, total; };
__host__ __device__ inline PtSmem pt_layout() {
  PtSmem s; uint32_t o = 0;
  s.stage = o; o += kPtStages * kPtTileBytes;
  s.aop = o; o += kPtGroups * kPtABytes;
  s.bars = o; o += 256;
  s.slot = o; o += 128;
  s.consts = o; o += 4 * 16 * 4;
  s.total = o;
  return s;
}

__global__ void __launch_bounds__(kPtThreads, 1) ffbsi_pair_tc_kernel(PairTcArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  griddep_entry();
  const PtSmem L = pt_layout();
  const int N = a.N, M = a.M, T = a.T, t = a.t, Dx = a.Dx;
  const bool boot = a.flags & PSVO_F_BOOTSTRAP;
  const bool last = t == T - 1;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int NT = (N + kPtCols - 1) / kPtCols;
  const long long ncand = (long long)N * M;
  const int rt_per_seq = (int)((ncand + 127) / 128);
  const int items_per_seq = (rt_per_seq + kPtGroups - 1) / kPtGroups;
  const int work = a.B * items_per_seq;

  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint64_t* full = bars;                          // [S]   B stage loaded        (1 arrival + tx bytes)
  uint64_t* empty = bars + kPtStages;             // [S]   B stage consumed      (one tcgen05.commit per MMA warp)
  uint64_t* dfull = bars + 2 * kPtStages;         // [g*2+buf] accumulator ready (tcgen05.commit)
  uint64_t* dempty = dfull + 2 * kPtGroups;       // [g*2+buf] accumulator read  (128 arrivals)
  uint64_t* afull = dempty + 2 * kPtGroups;       // [g]   A operand staged      (128 arrivals)
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + L.slot);
  float* C = reinterpret_cast<float*>(sm + L.consts);   // sp[16] | inv1[16] | inv2[16] | sf[16]

  if (tid < 16) {
    const int d = tid;
    const float s1 = a.sig[(last ? SIG_BQI : SIG_Q1I) * 32 + d], s2 = a.sig[SIG_BQ2 * 32 + d];
    C[16 + d] = 1.0f / s1; C[32 + d] = 1.0f / s2;
    C[d] = last ? s1 : 1.0f / (1.0f / s1 + 1.0f / s2);
    C[48 + d] = a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
  }
  if (tid == 32) {
    for (int s = 0; s < kPtStages; ++s) { tc::mbar_init(full + s, 1); tc::mbar_init(empty + s, 2); }
    for (int k = 0; k < 2 * kPtGroups; ++k) { tc::mbar_init(dfull + k, 1); tc::mbar_init(dempty + k, 128); }
    for (int k = 0; k < kPtGroups; ++k) tc::mbar_init(afull + k, 128);
  }
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;

  if (warp == kPtGroups * 4 + 2) {
    // ================================ TMA producer ================================
    if (tc::elect_one()) {
      uint32_t q = 0;
      for (int wi = blockIdx.x; wi < work; wi += gridDim.x) {
        const int b = wi / items_per_seq;
        const unsigned char* src = a.bimg + (size_t)b * NT * kPtTileBytes;
        for (int ct = 0; ct < NT; ++ct, ++q) {
          const uint32_t s = q % kPtStages, ph = (q / kPtStages) & 1u;
          tc::mbar_wait(empty + s, ph ^ 1u);
          tc::mbar_arrive_expect_tx(full + s, kPtTileBytes);
          tc::bulk_g2s(sm + L.stage + s * kPtTileBytes, src + (size_t)ct * kPtTileBytes, kPtTileBytes, full + s);
        }
      }
    }
    __syncwarp();
  } else if (warp >= kPtGroups * 4) {
    // ================================ MMA issuers: warp 16 serves groups 0-1, warp 17 groups 2-3 ================================
    if (tc::elect_one()) {
      const int g0 = (warp - kPtGroups * 4) * 2;
      const uint32_t idesc = tc::make_idesc_f16(128, kPtCols);
      uint64_t adesc_hi[2], adesc_lo[2];
      for (int k = 0; k < 2; ++k) {
        adesc_hi[k] = tc::make_desc(tc::smem_u32(sm + L.aop + (g0 + k) * kPtABytes), 2048, 128);
        adesc_lo[k] = tc::make_desc(tc::smem_u32(sm + L.aop + (g0 + k) * kPtABytes + 8192), 2048, 128);
      }
      uint32_t q = 0, it = 0;
      for (int wi = blockIdx.x; wi < work; wi += gridDim.x, ++it) {
        tc::mbar_wait(afull + g0, it & 1u);
        tc::mbar_wait(afull + g0 + 1, it & 1u);
        tc::fence_after_sync();
        for (int ct = 0; ct < NT; ++ct, ++q) {
          const uint32_t s = q % kPtStages;
          tc::mbar_wait(full + s, (q / kPtStages) & 1u);
          tc::fence_after_sync();
          const uint64_t bhi = tc::make_desc(tc::smem_u32(sm + L.stage + s * kPtTileBytes), kPtKg, 128);
          const uint64_t blo = tc::make_desc(tc::smem_u32(sm + L.stage + s * kPtTileBytes + 4 * kPtKg), kPtKg, 128);
          const uint32_t buf = q & 1u;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int g = g0 + k;
            tc::mbar_wait(dempty + g * 2 + buf, ((q >> 1) & 1u) ^ 1u);
            tc::fence_after_sync();
            const uint32_t d = tmem + (uint32_t)(g * 2 + buf) * (uint32_t)kPtCols;
            tc::mma_bf16(d, adesc_hi[k], bhi, idesc, false);                                          // hi * hi
            tc::mma_bf16(d, adesc_hi[k] + (4096 >> 4), bhi + ((2 * kPtKg) >> 4), idesc, true);          // 256 * (c_hi + c_lo)
            tc::mma_bf16(d, adesc_hi[k], blo, idesc, true);                                           // hi * lo
            tc::mma_bf16(d, adesc_lo[k], bhi, idesc, true);                                           // lo * hi
            tc::mma_commit(dfull + g * 2 + buf);
          }
          tc::mma_commit(empty + s);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ compute groups ================================
    const int g = warp >> 2;
    const int gt = tid & 127;
    const float* sp = C; const float* inv1 = C + 16; const float* inv2 = C + 32; const float* sf = C + 48;
    unsigned char* pa = sm + L.aop + g * kPtABytes + gt * 16;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const float kexp = 1.4426950408889634f / (kPtScale * kPtScale);
    uint32_t q = 0;
    for (int wi = blockIdx.x; wi < work; wi += gridDim.x) {
      const int b = wi / items_per_seq;
      const int rt = (wi - b * items_per_seq) * kPtGroups + g;
      const long long c = (long long)rt * 128 + gt;
      const bool live = rt < rt_per_seq && c < ncand;
      const int i = live ? (int)(c / M) : 0, m = live ? (int)(c - (long long)i * M) : 0;
      const size_t bt = (size_t)b * T + t;

      // ---- candidate x~ = mean + sp * eps (same arithmetic as ffbsi_step_kernel), scaled by 1 / s_f
      float xs[16];
      float rr = 0.0f;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (4 * cc < Dx) {
          if (a.eps_bw != nullptr) {
            const float* src = a.eps_bw + (((bt * M + m) * N + i) * (size_t)Dx);
#pragma unroll
            for (int k = 0; k < 4; ++k) if (4 * cc + k < Dx) e[k] = src[4 * cc + k];
          } else {
            const uint64_t gidx = ((((uint64_t)(a.seq_offset + b) * T + t) * M + m) * N) + i;
            normal4(philox4x32_10(a.seed, gidx, cc, 2u), e);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int d = 4 * cc + k;
          float xv = 0.0f;
          if (d < Dx && live) {
            float mean;
            if (last) mean = a.mu_init_all[bt * Dx + d];
            else {
              const float m1 = a.mu1i[((size_t)b * N + i) * Dx + d], m2 = a.m2_all[bt * Dx + d];
              mean = sp[d] * (inv1[d] * m1 + inv2[d] * m2);
            }
            xv = (mean + sp[d] * e[k]) / sf[d];
          }
          xs[d] = xv;
          rr = fmaf(xv, xv, rr);
        }
      }
      rr *= -0.5f;
#pragma unroll
      for (int kg = 0; kg < 2; ++kg) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = xs[kg * 8 + k] * kPtScale;
        uint4 hi, lo;
        tc::split8_f16(v, hi, lo);
        *reinterpret_cast<uint4*>(pa + kg * 2048) = hi;
        *reinterpret_cast<uint4*>(pa + 8192 + kg * 2048) = lo;
      }
      {
        const __half2 one2 = __floats2half2_rn(kPtScale * kPtScale, kPtScale * kPtScale);
        const uint32_t o = live ? *reinterpret_cast<const uint32_t*>(&one2) : 0u;
        *reinterpret_cast<uint4*>(pa + 2 * 2048) = make_uint4(o, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(pa + 3 * 2048) = make_uint4(0u, 0u, 0u, 0u);
      }
      tc::fence_proxy_async();
      tc::fence_before_sync();
      tc::mbar_arrive(afull + g);

      // ---- online log-sum-exp over the column tiles (accumulator units: 256 * natural log).  The TMEM read of the next
      //      32 columns is in flight while the current 32 are folded in (two register buffers); the accumulator is
      //      handed back to the MMA warp as soon as its last columns are in registers.
      float mD = -INFINITY, sum = 0.0f;
      auto fold = [&](const uint32_t (&r)[32]) {
        float cm = __uint_as_float(r[0]);
#pragma unroll
        for (int k = 1; k < 32; ++k) cm = fmaxf(cm, __uint_as_float(r[k]));
        if (cm > mD) { sum *= fast_ex2((mD - cm) * kexp); mD = cm; }
        const float off = -mD * kexp;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          s0 += fast_ex2(fmaf(__uint_as_float(r[k]), kexp, off));
          s1 += fast_ex2(fmaf(__uint_as_float(r[k + 1]), kexp, off));
          s2 += fast_ex2(fmaf(__uint_as_float(r[k + 2]), kexp, off));
          s3 += poly_ex2(fmaf(__uint_as_float(r[k + 3]), kexp, off));
        }
        sum += (s0 + s1) + (s2 + s3);
      };
      uint32_t va[32], vb[32];
      for (int ct = 0; ct < NT; ++ct, ++q) {
        const uint32_t buf = q & 1u;
        tc::mbar_wait(dfull + g * 2 + buf, (q >> 1) & 1u);
        tc::fence_after_sync();
        const uint32_t td = tmem + (uint32_t)(g * 2 + buf) * (uint32_t)kPtCols + lane_base;
        tc::tmem_ld32_issue(td, va);
        tc::tmem_ld32_issue(td + 32, vb);
        tc::tmem_ld32_wait(va);
        tc::tmem_ld32_wait(vb);
        tc::fence_before_sync();
        tc::mbar_arrive(dempty + g * 2 + buf);        // all 64 columns are in registers: the MMA warp may overwrite them
        fold(va);
        fold(vb);
      }
      if (live) a.candP[(size_t)b * ncand + c] = (mD * (1.0f / (kPtScale * kPtScale)) + logf(sum)) + rr;
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

bool ffbsi_pair_tc_supported(int N, int M, int B, int Dx) {
  static const bool disabled = getenv("PSVO_DISABLE_TC") != nullptr && getenv("PSVO_DISABLE_TC")[0] == '1';
  if (disabled || Dx > 16) return false;
  // one CTA item = 256 candidate rows x all N columns; below a machine-filling amount of work the SIMT loop wins
  return N >= 128 && (long long)B * N * M >= 2048;
}

size_t ffbsi_pair_tc_image_bytes(int B, int N) { return (size_t)B * ((N + kPtCols - 1) / kPtCols) * kPtTileBytes; }

int launch_ffbsi_pair_prep_all(PairTcArgs a, cudaStream_t st) {
  PSVO_REQUIRE(a.T >= 2 && a.T - 1 <= 65535, "pairwise B images of all steps: T out of range");
  a.prep_all = 1;
  ProfScope ps_(PK_FFBSI, st);
  PSVO_CUDA(launch_pdl(ffbsi_pair_prep_kernel, dim3((a.N + 127) / 128, a.B, a.T - 1), dim3(128), 0, st, pdl_enabled(), a));
  PSVO_LAUNCH_CHECK();
  return 0;
}

int launch_ffbsi_pair_tc(const PairTcArgs& a, cudaStream_t st) {
  if (!a.prep_done) {
    ProfScope ps_(PK_FFBSI, st);
    PSVO_CUDA(launch_pdl(ffbsi_pair_prep_kernel, dim3((a.N + 127) / 128, a.B), dim3(128), 0, st, pdl_enabled(), a));
    PSVO_LAUNCH_CHECK();
  }
  const PtSmem L = pt_layout();
  const size_t smem = L.total + 128;
  PSVO_DYN_SMEM(ffbsi_pair_tc_kernel, smem);
  const long long ncand = (long long)a.N * a.M;
  const long long work = (long long)a.B * (((ncand + 127) / 128 + kPtGroups - 1) / kPtGroups);
  const int grid = (int)(work < num_sms() ? work : num_sms());
  ProfScope ps_(PK_FFBSI, st);
  PSVO_CUDA(launch_pdl(ffbsi_pair_tc_kernel, dim3(grid), dim3(kPtThreads), smem, st, pdl_enabled(), a));
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo
