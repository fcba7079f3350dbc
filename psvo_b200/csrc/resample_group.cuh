// resample_group.cuh -- normalise + resample ONE sequence with ONE 128-thread compute group (the resampling step inside the
// persistent tensor-core filter kernel, smc_propagate_tc.cu).
//
// Reference: SVO.get_resample_idx (src/SMC/SVO.py:266-300) + the per-step logsumexp of compute_log_ZSMC (:302-311).
// Same rule, bit for bit, as resample_body (smc_forward.cu): deterministic exp -> 32.32 fixed-point weights -> integer
// inclusive prefix sum -> guide table over the uniform's 24-bit range -> integer upper bound against floor(u * total).
// Integer sums do not depend on how they are grouped, so this decomposition (every thread owns a CONTIGUOUS run of particles:
// a serial running sum per thread, one shuffle scan over the 32 thread totals of a warp, four warp totals through shared
// memory) gives the CDF of the 1024-thread kernel with a quarter of its instructions (~16 k warp-instructions for 4096
// particles: no per-32-particle shuffle scans, six group barriers).
//
// Shared-memory arrays carry one pad entry per thread chunk (index i lives at i + (i >> shift)), so that both the
// thread-contiguous walks (stride chunk + 1 entries between lanes: odd) and the thread-strided ones are bank-conflict free.
#pragma once
#include "smc_kernels.h"

namespace psvo {

constexpr int kRgThreads = 128;

// log2 of the thread chunk: smallest power of two >= ceil(n / nt)
__host__ __device__ inline int rg_chunk_shift(int n, int nt = kRgThreads) {
  const int cl = (n + nt - 1) / nt;
  int s = 0;
  while ((1 << s) < cl) ++s;
  return s;
}
// guide slices per particle (a power of two): more slices = narrower brackets for the inverse-CDF search, a longer table to build
#ifndef PSVO_GUIDE_FACTOR
#define PSVO_GUIDE_FACTOR 1
#endif
__host__ __device__ inline int rg_guide_bits(int N, int nt = kRgThreads) {
  int kbits = 0;
  while ((1 << kbits) < nt) ++kbits;                 // K >= nt: one guide slice per thread at least
  while ((1 << kbits) < N * PSVO_GUIDE_FACTOR) ++kbits;
  return kbits < 24 ? kbits : 24;
}
// bytes of the region: [lock + flags + cross-warp scratch: 1 KB] | CDF (u64) | log weights (f32) | guide table (i32)
__host__ __device__ inline size_t rg_smem_bytes(int N, int nt = kRgThreads) {
  const size_t K = (size_t)1 << rg_guide_bits(N, nt);
  return 1024 + ((size_t)N + nt + 8) * 8 + ((size_t)N + nt + 8) * 4 + (K + nt + 8) * 4;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel_gpu(int* p, int v) {
  int old;
  asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
struct RgGroupBar {            // named barrier of one 128-thread compute group
  int id;
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
};
struct RgBlockBar {            // the whole CTA
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
__device__ __forceinline__ void rg_bar(int bar_id) { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); }

// sequence b, time t; NT threads (a multiple of 32; g = thread index, bar = their barrier).  Writes anc[b][t][:], lse[b][t],
// accumulates logZ_b[b] (and logWbar when resampling is off); sAnc (optional, shared memory) also receives the ancestors
// (sAnc[n - c_lo]) for a gather by the caller.  [c_lo, c_hi): the children this caller serves (a CTA of a cluster rebuilds the
// whole CDF but resamples only its own range; write_stats: only one of them stores lse / logZ_b).  LDCG: log_Ws[b][t][:] was written by other SMs during this launch -> read through L2.
// region: rg_smem_bytes(N, NT) bytes; the caller owns the first 64 bytes (lock, flags).
template <bool LDCG, typename Bar>
__device__ __forceinline__ void resample_lean(const ResampleArgs& a, int b, int t, unsigned char* region, int g, const int NT, Bar bar,
                                              int* sAnc = nullptr, int c_lo = 0, int c_hi = 0x7fffffff, bool write_stats = true) {
  const int NW = NT >> 5;
  const int N = a.N;
  const int lane = g & 31, w = g >> 5;
  const int sh = rg_chunk_shift(N, NT), cl = 1 << sh;
  const int kbits = rg_guide_bits(N, NT), K = 1 << kbits;
  int ntb = 5;                                             // the guide table is scanned by the first VT = 2^ntb <= NT threads
  while ((2 << ntb) <= NT) ++ntb;
  const int VT = 1 << ntb, gsh = kbits - ntb, gl = 1 << gsh, gs = 24 - kbits;
  // region header (1 KB): words 0..15 belong to the caller | 64: NW warp maxima | 256: NW warp totals of the weights, then the
  // grand total | 768: NW warp totals of the histogram
  float* sred = reinterpret_cast<float*>(region + 64);
  unsigned long long* sws = reinterpret_cast<unsigned long long*>(region + 256);
  int* gws = reinterpret_cast<int*>(region + 768);
  unsigned long long* sC = reinterpret_cast<unsigned long long*>(region + 1024);
  float* sLW = reinterpret_cast<float*>(sC + N + NT + 8);
  int* sG = reinterpret_cast<int*>(sLW + N + NT + 8);
  const size_t bt = (size_t)b * a.T + t;
  const float* lw = a.log_Ws + bt * N;
  const bool resample = a.flags & PSVO_F_RESAMPLE;
  const bool guided = resample && a.forced_anc == nullptr;

  // ---- coalesced load of the log weights, running max; the guide table is cleared on the way
  float mx = -INFINITY;
  if ((N & 3) == 0) {
    for (int i = 4 * g; i < N; i += 4 * NT) {
      const float4 v = LDCG ? __ldcg(reinterpret_cast<const float4*>(lw + i)) : *reinterpret_cast<const float4*>(lw + i);
      sLW[i + (i >> sh)] = v.x; sLW[i + 1 + ((i + 1) >> sh)] = v.y; sLW[i + 2 + ((i + 2) >> sh)] = v.z; sLW[i + 3 + ((i + 3) >> sh)] = v.w;
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    }
  } else {
    for (int i = g; i < N; i += NT) { const float v = LDCG ? __ldcg(lw + i) : lw[i]; sLW[i + (i >> sh)] = v; mx = fmaxf(mx, v); }
  }
  if (guided) for (int i = g; i < K + NT + 8; i += NT) sG[i] = 0;
  mx = warp_max(mx);
  if (lane == 0) sred[w] = mx;
  bar();
  mx = warp_max(lane < NW ? sred[lane] : -INFINITY);

  // ---- integer CDF: serial running sum over this thread's contiguous run, shuffle scan of the thread totals, warp totals
  //      through shared memory, offsets added back
  const int lo = min(g * cl, N), hi = min(lo + cl, N);
  unsigned long long run = 0ull;
#pragma unroll 4
  for (int i = lo; i < hi; ++i) { run += fixed_weight(sLW[i + g], mx); sC[i + g] = run; }
  unsigned long long inc = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long nb = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += nb;
  }
  if (lane == 31) sws[w] = inc;
  bar();
  unsigned long long total, off = inc - run;
  {
    // every warp scans the NW warp totals itself (one shuffle scan; no extra barrier)
    const unsigned long long wt = lane < NW ? sws[lane] : 0ull;
    unsigned long long winc = wt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long nb = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += nb;
    }
    total = __shfl_sync(0xffffffffu, winc, 31);
    off += __shfl_sync(0xffffffffu, winc - wt, w);
  }
  if (off != 0ull) {
#pragma unroll 4
    for (int i = lo; i < hi; ++i) sC[i + g] += off;
  }
  bar();

  // logsumexp from the same integer total (identical to resample_body)
  const float lse = mx + (logf((float)total) - 32.0f * kLn2);
  if (g == 0 && write_stats) {
    if (a.lse != nullptr) a.lse[bt] = lse;
    if (a.logZ_b != nullptr) a.logZ_b[b] = (t == 0 ? 0.0f : a.logZ_b[b]) + lse;
  }

  if (guided) {
    // ---- guide table G[k] = #{j : C[j] <= B_k}: histogram of kf(j) = min{k : B_k >= C[j]}, then an inclusive scan
    const float scale = (float)K / (float)total;
    for (int j = g; j < N; j += NT) {
      const unsigned long long c = sC[j + (j >> sh)];
      int kf = (int)((float)c * scale);
      kf = kf < 0 ? 0 : (kf > K ? K : kf);
      while (kf > 0 && draw_threshold(total, (uint32_t)(kf - 1) << gs) >= c) --kf;
      while (kf < K && draw_threshold(total, (uint32_t)kf << gs) < c) ++kf;
      atomicAdd(&sG[kf + (kf >> gsh)], 1);
    }
    bar();
    const int g_lo = g * gl;
    int tot = 0, ginc = 0;
    if (g < VT) {
#pragma unroll 4
      for (int i = 0; i < gl; ++i) tot += sG[g_lo + i + g];
      ginc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, ginc, o);
        if (lane >= o) ginc += nb;
      }
      if (lane == 31) gws[w] = ginc;
    }
    bar();
    if (g < VT) {
      int grun = ginc - tot;
      const int wt = lane < (VT >> 5) ? gws[lane] : 0;
      int winc = wt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += nb;
      }
      grun += __shfl_sync(0xffffffffu, winc - wt, w);
#pragma unroll 4
      for (int i = 0; i < gl; ++i) { grun += sG[g_lo + i + g]; sG[g_lo + i + g] = grun; }
      if (g == 0) sG[K + (K >> gsh)] = N;                  // G[K]: every C[j] <= B_K
    }
    bar();
  }

  // ---- children: four CONSECUTIVE children per thread at a time (their uniforms are the four words of one Philox block)
  const bool systematic = a.flags & PSVO_F_SYSTEMATIC;
  const uint64_t gbase = ((uint64_t)(a.seq_offset + b) * a.T + t) * N;
  c_hi = c_hi < N ? c_hi : N;
  const bool vec_store = ((N | c_lo) & 3) == 0;
  for (int n0 = c_lo + 4 * g; n0 < c_hi; n0 += 4 * NT) {
    unsigned long long thr[4];
    int anc[4];
    bool search[4];
    int kb[4] = {0, 0, 0, 0};
    const bool draw = guided && a.u == nullptr;
    const bool block4 = draw && !systematic && ((gbase + (uint64_t)n0) & 3ull) == 0ull;
    Philox4 blk;
    blk.v[0] = blk.v[1] = blk.v[2] = blk.v[3] = 0u;
    if (block4) blk = philox4x32_10(a.seed, (gbase + (uint64_t)n0) >> 2, 0u, a.stream_id);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int n = n0 + k;
      anc[k] = n; search[k] = false; thr[k] = 0ull;
      if (n < c_hi && resample) {
        if (a.forced_anc != nullptr) {
          const int f = a.forced_anc[bt * N + n];
          anc[k] = f < 0 ? 0 : (f >= N ? N - 1 : f);
        } else {
          const int nu = systematic ? 0 : n;
          uint32_t u24;
          if (a.u != nullptr) u24 = u24_from_float(a.u[bt * N + nu]);
          else if (block4) u24 = blk.v[k] >> 8;
          else u24 = uniform_bits(a.seed, gbase + (uint64_t)nu, a.stream_id) >> 8;
          if (systematic) u24 = (uint32_t)((((uint64_t)n << 24) + u24) / (uint64_t)N);
          thr[k] = draw_threshold(total, u24);
          search[k] = true;
          kb[k] = min((int)(u24 >> gs), K - 1);
        }
      }
    }
    int pos[4], hi4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      pos[k] = search[k] ? sG[kb[k] + (kb[k] >> gsh)] : 0;
      hi4[k] = search[k] ? sG[kb[k] + 1 + ((kb[k] + 1) >> gsh)] : 0;
    }
    bool any = true;
    while (any) {                                // the four bracketed searches advance together
      any = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (pos[k] < hi4[k]) {
          const int mid = (pos[k] + hi4[k]) >> 1;
          if (sC[mid + (mid >> sh)] <= thr[k]) pos[k] = mid + 1; else hi4[k] = mid;
          any = true;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (search[k]) anc[k] = pos[k] < N ? pos[k] : N - 1;
      if (!resample && n0 + k < c_hi) a.logWbar[(size_t)b * N + n0 + k] = sLW[n0 + k + ((n0 + k) >> sh)] - lse;   // SVO.py:102,156
    }
    if (vec_store && n0 + 4 <= c_hi) {
      *reinterpret_cast<int4*>(a.anc + bt * N + n0) = make_int4(anc[0], anc[1], anc[2], anc[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (n0 + k < c_hi) a.anc[bt * N + n0 + k] = anc[k];
    }
    if (sAnc != nullptr) {
#pragma unroll
      for (int k = 0; k < 4; ++k) if (n0 + k < c_hi) sAnc[n0 + k - c_lo] = anc[k];
    }
  }
}

// the 128-thread compute-group form used inside the persistent tensor-core filter kernel
__device__ __forceinline__ void resample_group(const ResampleArgs& a, int b, int t, unsigned char* region, int g, int bar_id) {
  resample_lean<true>(a, b, t, region, g, kRgThreads, RgGroupBar{bar_id});
}

}  // namespace psvo
