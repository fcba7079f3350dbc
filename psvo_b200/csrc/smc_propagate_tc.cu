// smc_propagate_tc.cu -- forward-filter "propagate + weight" step with the per-particle MLPs on sm_100a tensor
// cores (tcgen05.mma, TMEM accumulators).  Same contract as smc_propagate_kernel (smc_forward.cu).
//
// Reference: SVO.SMC while_body (src/SMC/SVO.py:123-165) incl. sample_from_2_dist (:182-197,228-232); the two
// per-particle networks are MLP q1 (= f under use_bootstrap) and MLP g (src/transformation/MLP.py:48-54).
//
// Shape of the work (BASELINE config 5: Dx=16, Dy=32, H=256): per particle-step 41 kFLOP of tiny-K GEMMs
//   [128 x 16] x [16 x 256] -> relu -> [128 x 256] x [256 x 16|32].
//
// CTA = 4 compute groups of 128 threads + 4 MMA-issuing warps (one per group).  A group owns one 128-particle tile
// at a time (TMEM lane == row == thread) and walks the hidden layer of each network in chunks of 32 units:
//     compute threads : wait L1(c) -> tcgen05.ld the D1 row -> scale + bias -> ReLU -> fp16 hi/lo split (registers)
//                       -> wait L2(c-1) -> 16-byte stores into the A2 operand -> fence.proxy.async ->
//                       mbarrier.arrive full[g]          (no block- or group-wide barrier in the loop)
//     MMA warp g      : wait full[g] -> L2(c) (2 K-steps x 2 MMAs) -> commit ; L1(c+2) (3 MMAs) -> commit
// D1 is double buffered in TMEM so L1 runs two chunks ahead of the epilogue, and the wait for L2(c-1) sits after the
// register work of chunk c, so the compute warps only ever stall when the tensor pipe is the bottleneck.
// Measured cost model (profiles/microbench): a K=16 tcgen05.mma occupies the tensor pipe for max(40, N/2) cycles
// whatever its size, so MMAs are made as wide as TMEM allows: the layer-2 weights are stacked [w_hi ; w_lo] along N,
// which turns hi*hi + hi*lo into ONE instruction (the two halves of the accumulator are added in the epilogue).
// All groups share the staged fp16 weights (80 KB at H = 256).
// Arithmetic ("fp16x3"): x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 mantissa bits); a*w is evaluated as
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation in TMEM; weights are pre-scaled by 2^8 (exact) so that
// their lo halves stay in fp16's normal range, and the scale is undone inside the bias FFMA.  Operands must stay
// below fp16's 65504 (a particle that large has diverged; it turns into NaN and flags its sequence via logZ_b).
#include <stdlib.h>
#include "mlp_tile.cuh"
#include "smc_kernels.h"
#include "tc_common.cuh"

namespace psvo {

#ifdef PSVO_TC_TRACE
// low-overhead timeline of CTA 0: thread 0 (compute group 0) and thread 512 (its MMA warp); the last launch wins
__device__ long long* g_tc_trace = nullptr;
__shared__ int s_tc_trace_n[2];
#define TC_TRACE_INIT() do { if (threadIdx.x < 2) s_tc_trace_n[threadIdx.x] = 0; } while (0)
#define TC_TRACE(tag) do { if (blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 512)) { \
    long long* tb_ = g_tc_trace; const int w_ = threadIdx.x == 512; const int i_ = s_tc_trace_n[w_]; \
    if (tb_ != nullptr && (unsigned)i_ < 4096u) { s_tc_trace_n[w_] = i_ + 1; tb_[2 * (w_ * 4096 + i_)] = (tag) + w_ * 100; tb_[2 * (w_ * 4096 + i_) + 1] = clock64(); } } } while (0)
extern "C" __attribute__((visibility("default"))) int psvo_debug_tc_trace(long long* buf) {
  cudaMemcpyToSymbol(g_tc_trace, &buf, sizeof(buf));
  return 0;
}
#else
#define TC_TRACE(tag) do { } while (0)
#define TC_TRACE_INIT() do { } while (0)
#endif

constexpr int kTcGroups = 4;
constexpr int kTcChunk = 32;           // hidden units per pipeline stage
constexpr int kTcThreads = kTcGroups * 128 + kTcGroups * 32;
constexpr int kTcBars = 8;             // mbarriers per group (5 used)
// exact power-of-two operand scales: W1 (and its bias) by 2^4, W2 by 2^8 -> their fp16 lo halves stay normal; the hidden
// activations travel in 16x units (|h| < 4094) and the layer-2 accumulator in 4096x units, undone in the output FFMA
constexpr float kW1Scale = 16.0f, kW2Scale = 256.0f, kWUnscale = 1.0f / (16.0f * 256.0f);

struct TcPropSmem {
  uint32_t w1q_hi, w1q_lo, w2q, w1g_hi, w1g_lo, w2g;   // fp16 weights (B operands); w2*: [K-group][hi rows | lo rows]
  uint32_t w1q_b, w1g_b;                               // layer-1 bias as a B operand [H x K=16]: k=0 b_hi, k=1 b_lo
  uint32_t a_ones;                                     // its A operand [128 x K=16]: k=0,1 = 1 (same for every row)
  uint32_t b1q, b2q, b1g, b2g, consts;                                        // fp32
  uint32_t a1_hi[kTcGroups], a1_lo[kTcGroups];        // layer-1 A operands; the layer-2 A operand lives in tensor memory
  uint32_t rowc;                                       // per group, two tile parities: [mu2 (16) | obs (32) | pad] fp32
  uint32_t bars, slot, total;
};

__host__ __device__ inline TcPropSmem tc_prop_layout(int H, int NPG) {
  TcPropSmem s;
  uint32_t o = 0;
#define PSVO_TAKE(bytes) (o += (((bytes) + 127u) & ~127u), o - (((bytes) + 127u) & ~127u))
  s.w1q_hi = PSVO_TAKE(H * 16 * 2); s.w1q_lo = PSVO_TAKE(H * 16 * 2);
  s.w2q = PSVO_TAKE(2 * 16 * H * 2);
  s.w1g_hi = PSVO_TAKE(H * 16 * 2); s.w1g_lo = PSVO_TAKE(H * 16 * 2);
  s.w2g = PSVO_TAKE(2 * NPG * H * 2);
  s.w1q_b = PSVO_TAKE(H * 16 * 2); s.w1g_b = PSVO_TAKE(H * 16 * 2);
  s.b1q = PSVO_TAKE(H * 4); s.b2q = PSVO_TAKE(16 * 4); s.b1g = PSVO_TAKE(H * 4); s.b2g = PSVO_TAKE(32 * 4);
  s.consts = PSVO_TAKE(160 * 4);
  for (int g = 0; g < kTcGroups; ++g) {
    s.a1_hi[g] = PSVO_TAKE(128 * 16 * 2); s.a1_lo[g] = PSVO_TAKE(128 * 16 * 2);
  }
  s.a_ones = PSVO_TAKE(128 * 16 * 2);
  s.rowc = PSVO_TAKE(kTcGroups * 2 * 64 * 4);
  s.bars = PSVO_TAKE(kTcBars * kTcGroups * 8); s.slot = PSVO_TAKE(16);   // per group: full[0], mD1[0], mD1[1], mL2, mDone, full[1]
#undef PSVO_TAKE
  s.total = o;
  return s;
}

struct TcNetDesc { uint64_t w1h, w1l, w1b, w2; uint32_t idesc2a, idesc2b; int NP; };

struct TcGroup {
  unsigned char *pa1h, *pa1l;      // this thread's 16-byte slot in the A1 operand (K-group 0)
  uint64_t *full, *mD1, *mDone;    // full[0], full[5]: operand staged (128 arrivals, alternating) ; mD1: L1 into D1 done ; mDone: last L2 done
  uint32_t par;                    // parities of the next completions: bit 0 = mD1, bit 2 = mDone
  uint32_t k;                      // arrivals so far: arrival k goes to full[(k & 1) * 5]
  uint32_t tD1, tD2;               // TMEM addresses incl. this warp's lane offset: D1 = 64 columns (one chunk) ; D2 = 2*NP columns
};

__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

template <int BIT>
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t& par) {
  tc::mbar_wait(bar, (par >> BIT) & 1u);
  par ^= 1u << BIT;
}

// "operand staged" arrivals alternate between two mbarriers.  The compute threads may run one stage ahead of the MMA warp
// (arrival k only proves that stage k-2 was consumed: L1(c) is issued at stage c-1), and two completions of ONE mbarrier
// before the consumer's parity test would be indistinguishable from none.
__device__ __forceinline__ void tc_arrive(TcGroup& g) {
  tc::mbar_arrive(g.full + (g.k & 1u) * 5u);
  ++g.k;
}

// hidden-layer chunk (up to 64 hidden units = `halves` blocks of 32 TMEM columns, ONE layer-1 accumulator): per block
// TMEM row -> ReLU -> fp16 hi/lo -> back into the SAME 32 columns as the layer-2 A operand (columns 0-15: hi, two hidden
// units per column; 16-31: lo) -> arrive, so layer 2 of the first block runs while the second is being converted.
// The activations never touch shared memory: no st.shared, no generic->async proxy fence.
__device__ __forceinline__ void tc_chunk(TcGroup& g, int halves) {
  TC_TRACE(1);
  tc_wait<0>(g.mD1, g.par);                                               // L1 of the chunk complete
  tc::fence_after_sync();
  TC_TRACE(2);
#pragma unroll
  for (int hb = 0; hb < 2; ++hb) {
    if (hb < halves) {
      float v[32];
      tc::tmem_ld32(g.tD1 + hb * 32, v);
      uint32_t r[32];
#pragma unroll
      for (int kg = 0; kg < 4; ++kg) {
        float h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = v[kg * 8 + i];                  // D1 = 16 (W1^T x + b1): the bias rode in the MMA
        uint4 hi, lo;
        tc::split8_relu_f16(h, hi, lo);                                    // ReLU inside the packing conversion
        r[4 * kg] = hi.x; r[4 * kg + 1] = hi.y; r[4 * kg + 2] = hi.z; r[4 * kg + 3] = hi.w;
        r[16 + 4 * kg] = lo.x; r[16 + 4 * kg + 1] = lo.y; r[16 + 4 * kg + 2] = lo.z; r[16 + 4 * kg + 3] = lo.w;
      }
      tc::tmem_st32(g.tD1 + hb * 32, r);
      tc::tmem_st_wait();
      tc::fence_before_sync();
      tc_arrive(g);                                                        // A2 block staged in D1
    }
  }
  TC_TRACE(6);
}

// compute-group side of one network: x = this thread's input row.  On return the accumulator D2 (2*NP columns:
// [hi*hi + lo*hi | hi*lo]) is complete; (D2[:, :NP] + D2[:, NP:]) * kWUnscale = W2^T relu(W1^T x + b1).
__device__ __forceinline__ void tc_mlp_compute(TcGroup& g, const float (&x)[16], int nchunk) {
#pragma unroll
  for (int kg = 0; kg < 2; ++kg) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = x[kg * 8 + i];
    uint4 hi, lo;
    tc::split8_f16(v, hi, lo);
    *reinterpret_cast<uint4*>(g.pa1h + kg * 2048) = hi;
    *reinterpret_cast<uint4*>(g.pa1l + kg * 2048) = lo;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();           // also orders this thread's earlier tcgen05.ld of the previous accumulators
  tc_arrive(g);
  for (int c = 0; c < nchunk; c += 2) tc_chunk(g, nchunk - c < 2 ? nchunk - c : 2);   // nchunk counts blocks of 32 hidden units
  tc_wait<2>(g.mDone, g.par);                                              // L2(last) complete
  tc::fence_after_sync();
}

// MMA-warp side (one elected lane): L1 of the chunk starting at block c (c even; `halves` blocks of 32 hidden units) -> D1.
// One N = 64 instruction costs 48 tensor-pipe cycles against 2 x 40 for two N = 32 ones (profiles/microbench).
__device__ __forceinline__ void tc_issue_l1(const TcNetDesc& nd, uint64_t a1h, uint64_t a1l, uint64_t aones, uint32_t tD1, int c, int halves,
                                            uint64_t* mD1) {
  const uint32_t idesc1 = tc::make_idesc_f16(128, kTcChunk * halves);
  const uint64_t off = (uint64_t)(c * (kTcChunk * 16 / 16));
  tc::mma_bf16(tD1, a1h, nd.w1h + off, idesc1, false);
  tc::mma_bf16(tD1, a1h, nd.w1l + off, idesc1, true);
  tc::mma_bf16(tD1, a1l, nd.w1h + off, idesc1, true);
  tc::mma_bf16(tD1, aones, nd.w1b + off, idesc1, true);                    // + 16 b1 (hi + lo)
  tc::mma_commit(mD1);
}
// L2 of block c (32 hidden units), A operand in TMEM (D1 columns [32 (c & 1), +32)): per K-step  a_hi x [w_hi ; w_lo] (N = 2 NP)
// and  a_lo x w_hi (N = NP)
__device__ __forceinline__ void tc_issue_l2(const TcNetDesc& nd, uint32_t tD1, uint32_t tD2, int c) {
  const uint32_t a = tD1 + (c & 1) * 32;
#pragma unroll
  for (int ks = 0; ks < kTcChunk / 16; ++ks) {
    const uint64_t b = nd.w2 + (uint64_t)((uint32_t)(c * (kTcChunk / 8) + 2 * ks) * (uint32_t)(2 * nd.NP));   // (2*NP*16 B per K-group) >> 4
    tc::mma_f16_ts(tD2, a + 8 * ks, b, nd.idesc2a, (c > 0) || (ks > 0));
    tc::mma_f16_ts(tD2, a + 16 + 8 * ks, b, nd.idesc2b, true);
  }
}

// MMA-issuing warp of group `grp` (one elected lane): for every tile of the group and every network (first_net..1) the
// schedule  L1(chunk 0) | L2(block c) per staged block, and after the chunk's last block L1(next chunk)  driven by the
// group's "operand staged" barriers.  L1 of the next chunk overwrites the columns the chunk's L2s read their A operand from,
// so it is issued only after they have completed (mL2); the other three groups keep the tensor pipe busy meanwhile.
__device__ __forceinline__ void tc_mma_warp(unsigned char* sm, const TcPropSmem& L, uint32_t tmem, uint64_t* bars, int grp, int lane,
                                            int work, int nchunk, int first_net, const TcNetDesc& nq, const TcNetDesc& ng) {
  (void)lane;
  uint64_t* full = bars + kTcBars * grp; uint64_t* mD1 = full + 1; uint64_t* mL2 = full + 3; uint64_t* mDone = full + 4;
  const uint64_t a1h = tc::make_desc(tc::smem_u32(sm + L.a1_hi[grp]), 2048, 128), a1l = tc::make_desc(tc::smem_u32(sm + L.a1_lo[grp]), 2048, 128);
  const uint64_t aones = tc::make_desc(tc::smem_u32(sm + L.a_ones), 2048, 128);
  const uint32_t tD1 = tmem + grp * 128, tD2 = tmem + grp * 128 + 64;
  uint32_t k = 0, ph2 = 0;
  for (int wi = blockIdx.x * kTcGroups + grp; wi < work; wi += gridDim.x * kTcGroups) {
    for (int net = first_net; net < 2; ++net) {
      const TcNetDesc& nd = net ? ng : nq;
      for (int c = -1; c < nchunk; ++c) {                                  // c = -1: the network's input operand A1 is staged
        TC_TRACE(10);
        tc::mbar_wait(full + (k & 1u) * 5u, (k >> 1) & 1u); ++k;
        tc::fence_after_sync();
        TC_TRACE(11);
        if (tc::elect_one()) {
          if (c >= 0) tc_issue_l2(nd, tD1, tD2, c);
          const bool chunk_done = c < 0 || (c & 1) || c + 1 == nchunk;    // last block of its chunk
          if (chunk_done) {
            const int cn = c < 0 ? 0 : ((c | 1) + 1);                      // first block of the next chunk
            if (cn < nchunk) {
              if (c >= 0) {
                tc::mma_commit(mL2);
                TC_TRACE(13);
                tc::mbar_wait(mL2, ph2); ph2 ^= 1;                          // the chunk's A operand consumed -> D1 reusable
                tc::fence_after_sync();
              }
              tc_issue_l1(nd, a1h, a1l, aones, tD1, cn, nchunk - cn < 2 ? nchunk - cn : 2, mD1);
            } else {
              tc::mma_commit(mDone);
            }
          }
        }
        TC_TRACE(12);
        __syncwarp();
      }
    }
  }
}

// fp16 weight image (layout = the head of TcPropSmem), written once per forward pass
template <int NPG>
__global__ void __launch_bounds__(512) tc_stage_image_kernel(PropArgs a, unsigned char* __restrict__ img) {
  const int Dx = a.Dx, Dy = a.Dy, H = a.g.width[0];
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const int tid = threadIdx.x;
  for (uint32_t i = tid; i < L.consts / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(img)[i] = 0u;
  __syncthreads();
  tc::stage_weight_f16x2(a.q1.W[0], Dx, H, 16, H, kW1Scale, img + L.w1q_hi, img + L.w1q_lo);
  tc::stage_weight_f16x2_stacked(a.q1.W[1], H, Dx, H, 16, kW2Scale, img + L.w2q);
  tc::stage_weight_f16x2(a.g.W[0], Dx, H, 16, H, kW1Scale, img + L.w1g_hi, img + L.w1g_lo);
  tc::stage_weight_f16x2_stacked(a.g.W[1], H, Dy, H, NPG, kW2Scale, img + L.w2g);
  // layer-1 biases as B operands: row n of K-group 0 = [fp16(16 b), fp16(16 b - hi), 0 ...]; K-group 1 stays zero
  for (int n = tid; n < H; n += blockDim.x) {
    const float bq = __ldg(a.q1.b[0] + n) * kW1Scale, bg = __ldg(a.g.b[0] + n) * kW1Scale;
    const __half qh = __float2half_rn(bq), gh = __float2half_rn(bg);
    const __half2 q2 = __halves2half2(qh, __float2half_rn(bq - __half2float(qh)));
    const __half2 g2 = __halves2half2(gh, __float2half_rn(bg - __half2float(gh)));
    *reinterpret_cast<__half2*>(img + L.w1q_b + (size_t)n * 16) = q2;
    *reinterpret_cast<__half2*>(img + L.w1g_b + (size_t)n * 16) = g2;
  }
  float* b1q = reinterpret_cast<float*>(img + L.b1q); float* b2q = reinterpret_cast<float*>(img + L.b2q);
  float* b1g = reinterpret_cast<float*>(img + L.b1g); float* b2g = reinterpret_cast<float*>(img + L.b2g);
  for (int i = tid; i < H; i += blockDim.x) { b1q[i] = __ldg(a.q1.b[0] + i); b1g[i] = __ldg(a.g.b[0] + i); }
  for (int i = tid; i < 16; i += blockDim.x) b2q[i] = i < Dx ? __ldg(a.q1.b[1] + i) : 0.0f;
  for (int i = tid; i < 32; i += blockDim.x) b2g[i] = i < Dy ? __ldg(a.g.b[1] + i) : 0.0f;
}

template <int NPG>
__global__ void __launch_bounds__(kTcThreads, 1) smc_propagate_tc_kernel(PropArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int Dx = a.Dx, Dy = a.Dy, N = a.N, H = a.g.width[0];
  const bool two_q = a.flags & PSVO_F_TWO_Q, poisson = a.flags & PSVO_F_POISSON;
  const bool t0 = a.t == 0;
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform (kept in a uniform register)
  const int nchunk = H / kTcChunk;

  float* sb2q = reinterpret_cast<float*>(sm + L.b2q);
  float* sb2g = reinterpret_cast<float*>(sm + L.b2g);
  // consts: s1[16] | sp[16] | inv1[16] | inv2[16] | inv_sp[16] | inv_s1[16] | inv_sg[32] | ld_p ld_f ld_g
  float* C = reinterpret_cast<float*>(sm + L.consts);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);     // [g][kTcBars]: full, mD1[0], mD1[1], mL2, mDone
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + L.slot);

  // ---- weights (fp16 hi/lo, UMMA B layout, pre-scaled) and biases: one coalesced copy of the image that
  //      tc_stage_image_kernel prepared once for this forward pass
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.tc_image);
    uint4* dst = reinterpret_cast<uint4*>(sm);
    const int n16 = (int)(L.consts / 16);
#pragma unroll 4
    for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (tid < 16) {
    const int d = tid;
    const float s1 = a.sig[(t0 ? SIG_Q0 : SIG_Q1) * 32 + d], s2 = a.sig[SIG_Q2 * 32 + d];
    const float i1 = 1.0f / s1, i2 = 1.0f / s2;
    const float spv = two_q ? 1.0f / (i1 + i2) : s1;
    C[d] = s1; C[16 + d] = spv; C[32 + d] = i1; C[48 + d] = i2; C[64 + d] = 1.0f / spv; C[80 + d] = i1;
  }
  if (tid >= 32 && tid < 64) C[96 + tid - 32] = tid - 32 < Dy ? 1.0f / a.sig[SIG_G * 32 + tid - 32] : 0.0f;   // 0 beyond Dy: padded columns drop out
  if (tid < kTcBars * kTcGroups) tc::mbar_init(bars + tid, (tid % kTcBars) == 0 || (tid % kTcBars) == 5 ? 128 : 1);
  if (tid < 256) {                                                            // A operand of the bias MMA: every row [1, 1, 0 ...]
    const uint32_t one2 = 0x3C003C00u;                                        // half2(1, 1)
    *reinterpret_cast<uint4*>(sm + L.a_ones + tid * 16) = tid < 128 ? make_uint4(one2, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  TC_TRACE_INIT();
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    float lp = 0.f, lf = 0.f, lg = 0.f;
    for (int d = 0; d < Dx; ++d) { lp += logf(C[16 + d]); lf += logf(C[d]); }
    if (!poisson) for (int d = 0; d < Dy; ++d) lg += logf(a.sig[SIG_G * 32 + d]);
    C[128] = lp; C[129] = lf; C[130] = lg;
  }
  __syncthreads();
  const uint32_t tmem = *slot;
  const int tiles_per_seq = (N + 127) / 128;
  const int work = a.B * tiles_per_seq;

  TcNetDesc nq, ng;
  nq.w1h = tc::make_desc(tc::smem_u32(sm + L.w1q_hi), H * 16, 128); nq.w1l = tc::make_desc(tc::smem_u32(sm + L.w1q_lo), H * 16, 128);
  nq.w1b = tc::make_desc(tc::smem_u32(sm + L.w1q_b), H * 16, 128);
  nq.w2 = tc::make_desc(tc::smem_u32(sm + L.w2q), 2 * 16 * 16, 128);
  nq.idesc2a = tc::make_idesc_f16(128, 32); nq.idesc2b = tc::make_idesc_f16(128, 16); nq.NP = 16;
  ng.w1h = tc::make_desc(tc::smem_u32(sm + L.w1g_hi), H * 16, 128); ng.w1l = tc::make_desc(tc::smem_u32(sm + L.w1g_lo), H * 16, 128);
  ng.w1b = tc::make_desc(tc::smem_u32(sm + L.w1g_b), H * 16, 128);
  ng.w2 = tc::make_desc(tc::smem_u32(sm + L.w2g), 2 * NPG * 16, 128);
  ng.idesc2a = tc::make_idesc_f16(128, 2 * NPG); ng.idesc2b = tc::make_idesc_f16(128, NPG); ng.NP = NPG;

  if (warp >= kTcGroups * 4) {
    tc_mma_warp(sm, L, tmem, bars, warp - kTcGroups * 4, lane, work, nchunk, t0 ? 1 : 0, nq, ng);
  } else {
    // =================================== compute groups ===================================
    const int grp = warp >> 2;
    const float* s1 = C; const float* sp = C + 16; const float* inv1 = C + 32; const float* inv2 = C + 48;
    const float* inv_sp = C + 64; const float* inv_s1 = C + 80; const float* inv_sg = C + 96;
    const float ld_p = C[128], ld_f = C[129], ld_g = C[130];
    const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;
    const float neg_logN = -logf((float)N);
    (void)s1;
    TcGroup g;
    const int gt = tid & 127;
    g.full = bars + kTcBars * grp; g.mD1 = g.full + 1; g.mDone = g.full + 4; g.par = 0; g.k = 0;
    g.pa1h = sm + L.a1_hi[grp] + gt * 16; g.pa1l = sm + L.a1_lo[grp] + gt * 16;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    g.tD1 = tmem + grp * 128 + lane_base; g.tD2 = tmem + grp * 128 + 64 + lane_base;

    uint32_t tile_it = 0;
    for (int wi = blockIdx.x * kTcGroups + grp; wi < work; wi += gridDim.x * kTcGroups) {
      const int b = wi / tiles_per_seq;
      const int n = (wi - b * tiles_per_seq) * 128 + gt;
      const bool live = n < N;
      const size_t bt = (size_t)b * a.T + a.t;
      const size_t row = bt * N + (live ? n : 0);

      TC_TRACE(20);
      // ---- per-sequence rows every particle of the tile shares (second proposal mean, observation): one global load
      //      per thread into the group's double-buffered slot instead of 48 dependent broadcast loads per thread
      const float* rc = reinterpret_cast<const float*>(sm + L.rowc) + (grp * 2 + (tile_it & 1)) * 64;
      float rowv = 0.0f;                                                   // loaded now, parked in the slot after the Philox work
      if (two_q && gt < Dx) rowv = __ldg(a.mu2_all + bt * Dx + gt);
      if (gt >= 32 && gt < 32 + Dy) rowv = __ldg(a.obs + bt * Dy + gt - 32);
      ++tile_it;
      float x[16], mu1[16];
      // ---- d1 mean: MLP q1 on the resampled parent (t >= 1) or the per-sequence q0 mean (t = 0).  With the fused
      //      gather the parent row is Xs[t-1][anc[t-1][n]] (L2-resident: the previous launch wrote it) and is also
      //      copied to X_anc[t-1][n], the reference's log["Xs"] output (SVO.py:54)
      const float* xsrc = nullptr;
      if (!t0) {
        const size_t prow = (bt - 1) * N + (live ? n : 0);
        xsrc = a.anc_prev != nullptr ? a.Xs + ((bt - 1) * N + __ldg(a.anc_prev + prow)) * Dx : a.X_anc + prow * Dx;
      }
      // noise for this particle (the Philox work also covers the latency of the ancestor index load above)
      float e[16];
      if (a.eps != nullptr) {
        const float* src = a.eps + row * Dx;
#pragma unroll
        for (int d = 0; d < 16; ++d) e[d] = d < Dx ? src[d] : 0.0f;
      } else {
        const uint64_t gidx = ((uint64_t)(a.seq_offset + b) * a.T + a.t) * N + (live ? n : 0);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (4 * c < Dx) normal4(philox4x32_10(a.seed, gidx, c, 0u), &e[4 * c]);
          else { e[4 * c] = e[4 * c + 1] = e[4 * c + 2] = e[4 * c + 3] = 0.f; }
        }
      }
      if (!t0) {
        if (Dx == 16) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const float4 q = reinterpret_cast<const float4*>(xsrc)[v];
            x[4 * v] = q.x; x[4 * v + 1] = q.y; x[4 * v + 2] = q.z; x[4 * v + 3] = q.w;
          }
        } else {
#pragma unroll
          for (int d = 0; d < 16; ++d) x[d] = d < Dx ? xsrc[d] : 0.0f;
        }
      }
      if (a.eps == nullptr) {
        const uint64_t gidx = ((uint64_t)(a.seq_offset + b) * a.T + a.t) * N + (live ? n : 0);
#pragma unroll
        for (int c = 2; c < 4; ++c) {
          if (4 * c < Dx) normal4(philox4x32_10(a.seed, gidx, c, 0u), &e[4 * c]);
          else { e[4 * c] = e[4 * c + 1] = e[4 * c + 2] = e[4 * c + 3] = 0.f; }
        }
      }
      if (!t0 && a.anc_prev != nullptr && live) {
        float* dst = a.X_anc_w + ((bt - 1) * N + n) * Dx;
        if (Dx == 16) {
#pragma unroll
          for (int v = 0; v < 4; ++v) reinterpret_cast<float4*>(dst)[v] = make_float4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
        } else {
#pragma unroll
          for (int d = 0; d < 16; ++d) if (d < Dx) dst[d] = x[d];
        }
      }
      {
        float* rcw = const_cast<float*>(rc);
        if (gt < 16) rcw[gt] = rowv;
        if (gt >= 32 && gt < 64) rcw[16 + gt - 32] = rowv;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      }
      if (!t0) {
        tc_mlp_compute(g, x, nchunk);
        float mb[16];
        tc::tmem_ld16(g.tD2, mu1);
        tc::tmem_ld16(g.tD2 + 16, mb);
#pragma unroll
        for (int d = 0; d < 16; ++d) mu1[d] = fmaf(mu1[d] + mb[d], kWUnscale, sb2q[d]);
      } else {
#pragma unroll
        for (int d = 0; d < 16; ++d) mu1[d] = d < Dx ? __ldg(a.mu0 + (size_t)b * Dx + d) : 0.0f;
      }

      // ---- proposal (fusion), sample, q / f log densities   (f = d1: use_bootstrap, SVO.py:230)
      float qs = 0.f, fsum = 0.f;
      float mq[16];                                                        // proposal mean (fusion of d1 and d2, SVO.py:182-197)
      if (two_q) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 m2 = reinterpret_cast<const float4*>(rc)[v];       // rows beyond Dx hold 0
          mq[4 * v] = sp[4 * v] * (inv1[4 * v] * mu1[4 * v] + inv2[4 * v] * m2.x);
          mq[4 * v + 1] = sp[4 * v + 1] * (inv1[4 * v + 1] * mu1[4 * v + 1] + inv2[4 * v + 1] * m2.y);
          mq[4 * v + 2] = sp[4 * v + 2] * (inv1[4 * v + 2] * mu1[4 * v + 2] + inv2[4 * v + 2] * m2.z);
          mq[4 * v + 3] = sp[4 * v + 3] * (inv1[4 * v + 3] * mu1[4 * v + 3] + inv2[4 * v + 3] * m2.w);
        }
      } else {
#pragma unroll
        for (int d = 0; d < 16; ++d) mq[d] = mu1[d];
      }
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float m = mq[d];
        const float X = fmaf(sp[d], e[d], m);
        const float zq = (X - m) * inv_sp[d];
        const float zf = (X - mu1[d]) * inv_s1[d];
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

      TC_TRACE(21);
      // ---- emission
      tc_mlp_compute(g, x, nchunk);
      TC_TRACE(22);
      float gsum = 0.f;
#pragma unroll
      for (int c = 0; c < NPG; c += 16) {
        float mug[16], mgb[16];
        tc::tmem_ld16(g.tD2 + c, mug);
        tc::tmem_ld16(g.tD2 + NPG + c, mgb);
        if (!poisson) {                                                    // branch-free Gaussian emission; columns beyond Dy: y = b2 = 0
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int d = c + i;
            const float mu = fmaf(mug[i] + mgb[i], kWUnscale, sb2g[d]);
            const float z = (rc[16 + d] - mu) * inv_sg[d];
            gsum = fmaf(z, z, gsum);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int d = c + i;
            if (d < Dy) {
              const float mu = fmaf(mug[i] + mgb[i], kWUnscale, sb2g[d]);
              const float z = rc[16 + d] - (softplus(mu) + 1e-6f);
              gsum = fmaf(z, z, gsum);
            }
          }
        }
      }
      const float g_lp = -0.5f * gsum - ld_g - cy;
      float lwb = neg_logN;
      if (!t0 && a.logWbar != nullptr && live) lwb = a.logWbar[(size_t)b * N + n];
      if (live) a.log_Ws[bt * N + n] = ((f_lp + g_lp) - q_lp) + lwb;
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------
// FFBSi candidates: MLP f and MLP g of every candidate x~[m,i] on tensor cores  (PSVO.py:125,130 -- f.log_prob(bw_X_t,
// bw_X_tp1) and g.log_prob(bw_X_t, obs[:, t]) on M*N*B rows).  Same pipeline as smc_propagate_tc_kernel (the weight
// image holds f in the "q" slot); the candidate is recomputed from eps_bw / Philox exactly like ffbsi_step_kernel does.
// Outputs f_lp and g_lp per candidate, row index (b, i, m).
// ------------------------------------------------------------------------------------------------------
template <int NPG>
__global__ void __launch_bounds__(kTcThreads, 1) ffbsi_cand_tc_kernel(CandTcArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int Dx = a.Dx, Dy = a.Dy, N = a.N, M = a.M, T = a.T, t = a.t, H = a.H;
  const bool boot = a.flags & PSVO_F_BOOTSTRAP, poisson = a.flags & PSVO_F_POISSON;
  const bool last = t == T - 1;
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nchunk = H / kTcChunk;

  float* sb2q = reinterpret_cast<float*>(sm + L.b2q);
  float* sb2g = reinterpret_cast<float*>(sm + L.b2g);
  // consts: sp[16] | inv1[16] | inv2[16] | inv_sf[16] | inv_sg[32] | ld_f ld_g
  float* C = reinterpret_cast<float*>(sm + L.consts);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + L.slot);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.tc_image);
    uint4* dst = reinterpret_cast<uint4*>(sm);
    const int n16 = (int)(L.consts / 16);
#pragma unroll 4
    for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (tid < 16) {
    const int d = tid;
    const float s1 = a.sig[(last ? SIG_BQI : SIG_Q1I) * 32 + d], s2 = a.sig[SIG_BQ2 * 32 + d];
    C[16 + d] = 1.0f / s1; C[32 + d] = 1.0f / s2;
    C[d] = last ? s1 : 1.0f / (1.0f / s1 + 1.0f / s2);
    C[48 + d] = 1.0f / a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
  }
  if (tid >= 32 && tid < 64) C[64 + tid - 32] = tid - 32 < Dy ? 1.0f / a.sig[SIG_G * 32 + tid - 32] : 0.0f;   // 0 beyond Dy: padded columns drop out
  if (tid < kTcBars * kTcGroups) tc::mbar_init(bars + tid, (tid % kTcBars) == 0 || (tid % kTcBars) == 5 ? 128 : 1);
  if (tid < 256) {                                                            // A operand of the bias MMA: every row [1, 1, 0 ...]
    const uint32_t one2 = 0x3C003C00u;                                        // half2(1, 1)
    *reinterpret_cast<uint4*>(sm + L.a_ones + tid * 16) = tid < 128 ? make_uint4(one2, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (tid == 0) {
    float lf = 0.f, lg = 0.f;
    for (int d = 0; d < Dx; ++d) lf += logf(a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d]);
    if (!poisson) for (int d = 0; d < Dy; ++d) lg += logf(a.sig[SIG_G * 32 + d]);
    C[96] = lf; C[97] = lg;
  }
  __syncthreads();
  const uint32_t tmem = *slot;
  const long long ncand = (long long)N * M;
  const int tiles_per_seq = (int)((ncand + 127) / 128);
  const int work = a.B * tiles_per_seq;

  TcNetDesc nq, ng;
  nq.w1h = tc::make_desc(tc::smem_u32(sm + L.w1q_hi), H * 16, 128); nq.w1l = tc::make_desc(tc::smem_u32(sm + L.w1q_lo), H * 16, 128);
  nq.w1b = tc::make_desc(tc::smem_u32(sm + L.w1q_b), H * 16, 128);
  nq.w2 = tc::make_desc(tc::smem_u32(sm + L.w2q), 2 * 16 * 16, 128);
  nq.idesc2a = tc::make_idesc_f16(128, 32); nq.idesc2b = tc::make_idesc_f16(128, 16); nq.NP = 16;
  ng.w1h = tc::make_desc(tc::smem_u32(sm + L.w1g_hi), H * 16, 128); ng.w1l = tc::make_desc(tc::smem_u32(sm + L.w1g_lo), H * 16, 128);
  ng.w1b = tc::make_desc(tc::smem_u32(sm + L.w1g_b), H * 16, 128);
  ng.w2 = tc::make_desc(tc::smem_u32(sm + L.w2g), 2 * NPG * 16, 128);
  ng.idesc2a = tc::make_idesc_f16(128, 2 * NPG); ng.idesc2b = tc::make_idesc_f16(128, NPG); ng.NP = NPG;

  if (warp >= kTcGroups * 4) {
    tc_mma_warp(sm, L, tmem, bars, warp - kTcGroups * 4, lane, work, nchunk, last ? 1 : 0, nq, ng);
  } else {
    const int grp = warp >> 2;
    const float* sp = C; const float* inv1 = C + 16; const float* inv2 = C + 32; const float* inv_sf = C + 48;
    const float* inv_sg = C + 64;
    const float ld_f = C[96], ld_g = C[97];
    const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;
    TcGroup g;
    const int gt = tid & 127;
    g.full = bars + kTcBars * grp; g.mD1 = g.full + 1; g.mDone = g.full + 4; g.par = 0; g.k = 0;
    g.pa1h = sm + L.a1_hi[grp] + gt * 16; g.pa1l = sm + L.a1_lo[grp] + gt * 16;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    g.tD1 = tmem + grp * 128 + lane_base; g.tD2 = tmem + grp * 128 + 64 + lane_base;

    uint32_t tile_it = 0;
    for (int wi = blockIdx.x * kTcGroups + grp; wi < work; wi += gridDim.x * kTcGroups) {
      const int b = wi / tiles_per_seq;
      const long long c = (long long)(wi - b * tiles_per_seq) * 128 + gt;
      const bool live = c < ncand;
      const int i = live ? (int)(c / M) : 0, m = live ? (int)(c - (long long)i * M) : 0;
      const size_t bt = (size_t)b * T + t;

      // ---- per-sequence rows every candidate of the tile shares (proposal mean part, observation): one global load per
      //      thread into the group's double-buffered slot instead of dependent broadcast loads per element
      const float* rc = reinterpret_cast<const float*>(sm + L.rowc) + (grp * 2 + (tile_it & 1)) * 64;
      {
        float rowv = 0.0f;
        if (gt < Dx) rowv = last ? __ldg(a.mu_init_all + bt * Dx + gt) : __ldg(a.m2_all + bt * Dx + gt);
        if (gt >= 32 && gt < 32 + Dy) rowv = __ldg(a.obs + bt * Dy + gt - 32);
        float* rcw = const_cast<float*>(rc);
        if (gt < 16) rcw[gt] = rowv;
        if (gt >= 32 && gt < 64) rcw[16 + gt - 32] = rowv;
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      }
      ++tile_it;

      // ---- candidate x~ = mean + sp * eps  (same arithmetic as ffbsi_step_kernel)
      float x[16];
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
          if (d < Dx) {
            float mean;
            if (last) mean = rc[d];
            else {
              const float m1 = __ldg(a.mu1i + ((size_t)b * N + i) * Dx + d), m2 = rc[d];
              mean = sp[d] * (inv1[d] * m1 + inv2[d] * m2);
            }
            xv = mean + sp[d] * e[k];
          }
          x[d] = xv;
        }
      }

      // ---- f(x_{t+1} | x~)
      float f_lp = 0.0f;
      if (!last) {
        tc_mlp_compute(g, x, nchunk);
        float muf[16], mb[16];
        tc::tmem_ld16(g.tD2, muf);
        tc::tmem_ld16(g.tD2 + 16, mb);
        const float* xn = a.X_next + ((bt + 1) * N + i) * Dx;
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          if (d < Dx) {
            const float mu = fmaf(muf[d] + mb[d], kWUnscale, sb2q[d]);
            const float z = (xn[d] - mu) * inv_sf[d];
            s = fmaf(z, z, s);
          }
        }
        f_lp = -0.5f * s - ld_f - cx;
      }
      // ---- g(y_t | x~)
      tc_mlp_compute(g, x, nchunk);
      float gsum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < NPG; c0 += 16) {
        float mug[16], mgb[16];
        tc::tmem_ld16(g.tD2 + c0, mug);
        tc::tmem_ld16(g.tD2 + NPG + c0, mgb);
        if (!poisson) {                                                    // branch-free Gaussian emission; columns beyond Dy: y = b2 = 0
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int d = c0 + k;
            const float mu = fmaf(mug[k] + mgb[k], kWUnscale, sb2g[d]);
            const float z = (rc[16 + d] - mu) * inv_sg[d];
            gsum = fmaf(z, z, gsum);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int d = c0 + k;
            if (d < Dy) {
              const float mu = fmaf(mug[k] + mgb[k], kWUnscale, sb2g[d]);
              const float z = rc[16 + d] - (softplus(mu) + 1e-6f);
              gsum = fmaf(z, z, gsum);
            }
          }
        }
      }
      const float g_lp = -0.5f * gsum - ld_g - cy;
      if (live) {
        a.candF[(size_t)b * ncand + c] = f_lp;
        a.candG[(size_t)b * ncand + c] = g_lp;
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------
// Plain forward of the two per-particle networks for the reverse pass (smc_backward.cu recomputes the network outputs
// instead of storing them):  mu_q = MLP q1(xq rows), mu_g = MLP g(xg rows), same pipeline as smc_propagate_tc_kernel.
// ------------------------------------------------------------------------------------------------------
template <int NPG>
__global__ void __launch_bounds__(kTcThreads, 1) tc_fwd2_kernel(Fwd2TcArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int Dx = a.Dx, Dy = a.Dy, N = a.N, H = a.H;
  const bool with_q = a.xq != nullptr;
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nchunk = H / kTcChunk;
  float* sb2q = reinterpret_cast<float*>(sm + L.b2q);
  float* sb2g = reinterpret_cast<float*>(sm + L.b2g);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + L.slot);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.tc_image);
    uint4* dst = reinterpret_cast<uint4*>(sm);
    const int n16 = (int)(L.consts / 16);
#pragma unroll 4
    for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (tid < kTcBars * kTcGroups) tc::mbar_init(bars + tid, (tid % kTcBars) == 0 || (tid % kTcBars) == 5 ? 128 : 1);
  if (tid < 256) {
    const uint32_t one2 = 0x3C003C00u;
    *reinterpret_cast<uint4*>(sm + L.a_ones + tid * 16) = tid < 128 ? make_uint4(one2, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  const long long rows = (long long)a.B * N;
  const int work = (int)((rows + 127) / 128);

  TcNetDesc nq, ng;
  nq.w1h = tc::make_desc(tc::smem_u32(sm + L.w1q_hi), H * 16, 128); nq.w1l = tc::make_desc(tc::smem_u32(sm + L.w1q_lo), H * 16, 128);
  nq.w1b = tc::make_desc(tc::smem_u32(sm + L.w1q_b), H * 16, 128);
  nq.w2 = tc::make_desc(tc::smem_u32(sm + L.w2q), 2 * 16 * 16, 128);
  nq.idesc2a = tc::make_idesc_f16(128, 32); nq.idesc2b = tc::make_idesc_f16(128, 16); nq.NP = 16;
  ng.w1h = tc::make_desc(tc::smem_u32(sm + L.w1g_hi), H * 16, 128); ng.w1l = tc::make_desc(tc::smem_u32(sm + L.w1g_lo), H * 16, 128);
  ng.w1b = tc::make_desc(tc::smem_u32(sm + L.w1g_b), H * 16, 128);
  ng.w2 = tc::make_desc(tc::smem_u32(sm + L.w2g), 2 * NPG * 16, 128);
  ng.idesc2a = tc::make_idesc_f16(128, 2 * NPG); ng.idesc2b = tc::make_idesc_f16(128, NPG); ng.NP = NPG;

  if (warp >= kTcGroups * 4) {
    tc_mma_warp(sm, L, tmem, bars, warp - kTcGroups * 4, lane, work, nchunk, with_q ? 0 : 1, nq, ng);
  } else {
    const int grp = warp >> 2;
    TcGroup g;
    const int gt = tid & 127;
    g.full = bars + kTcBars * grp; g.mD1 = g.full + 1; g.mDone = g.full + 4; g.par = 0; g.k = 0;
    g.pa1h = sm + L.a1_hi[grp] + gt * 16; g.pa1l = sm + L.a1_lo[grp] + gt * 16;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    g.tD1 = tmem + grp * 128 + lane_base; g.tD2 = tmem + grp * 128 + 64 + lane_base;
    for (int wi = blockIdx.x * kTcGroups + grp; wi < work; wi += gridDim.x * kTcGroups) {
      const long long r = (long long)wi * 128 + gt;
      const bool live = r < rows;
      const int b = live ? (int)(r / N) : 0, n = live ? (int)(r - (long long)b * N) : 0;
      float x[16];
      if (with_q) {
        const float* src = a.xq + (long long)b * a.x_seq_stride + (long long)n * Dx;
#pragma unroll
        for (int d = 0; d < 16; ++d) x[d] = (live && d < Dx) ? src[d] : 0.0f;
        tc_mlp_compute(g, x, nchunk);
        float mu[16], mb[16];
        tc::tmem_ld16(g.tD2, mu);
        tc::tmem_ld16(g.tD2 + 16, mb);
        if (live) {
          float* dst = a.muq + r * Dx;
#pragma unroll
          for (int d = 0; d < 16; ++d) if (d < Dx) dst[d] = fmaf(mu[d] + mb[d], kWUnscale, sb2q[d]);
        }
      }
      {
        const float* src = a.xg + (long long)b * a.x_seq_stride + (long long)n * Dx;
#pragma unroll
        for (int d = 0; d < 16; ++d) x[d] = (live && d < Dx) ? src[d] : 0.0f;
        tc_mlp_compute(g, x, nchunk);
#pragma unroll
        for (int c0 = 0; c0 < NPG; c0 += 16) {
          float mu[16], mb[16];
          tc::tmem_ld16(g.tD2 + c0, mu);
          tc::tmem_ld16(g.tD2 + NPG + c0, mb);
          if (live) {
            float* dst = a.mug + r * Dy;
#pragma unroll
            for (int k = 0; k < 16; ++k) if (c0 + k < Dy) dst[c0 + k] = fmaf(mu[k] + mb[k], kWUnscale, sb2g[c0 + k]);
          }
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int launch_tc_fwd2(const Fwd2TcArgs& a, cudaStream_t st) {
  PSVO_REQUIRE(a.tc_image != nullptr && a.xg != nullptr && a.mug != nullptr, "tensor-core forward: null buffer");
  const TcPropSmem L = tc_prop_layout(a.H, 32);
  const size_t smem = L.total + 128;
  PSVO_REQUIRE(smem <= 227 * 1024, "tensor-core forward: %zu B of shared memory", smem);
  const long long tiles = ((long long)a.B * a.N + 127) / 128;
  long long grid = (tiles + kTcGroups - 1) / kTcGroups;
  if (grid > num_sms()) grid = num_sms();
  static bool cfg = false;
  if (!cfg) { PSVO_CUDA(cudaFuncSetAttribute(tc_fwd2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
  ProfScope ps_(PK_BWD, st);
  tc_fwd2_kernel<32><<<(int)grid, kTcThreads, smem, st>>>(a);
  PSVO_LAUNCH_CHECK();
  return 0;
}

// shapes the forward tensor-core pipeline handles (one hidden layer of 32..256 units in both networks)
bool tc_pipeline_supported(const psvo_mlp& q, const psvo_mlp& g, int Dx, int Dy) {
  if (q.n_hidden != 1 || g.n_hidden != 1) return false;
  const int H = g.width[0];
  return q.width[0] == H && H >= 32 && H <= 256 && (H % 32) == 0 && Dx <= 16 && Dy <= 32;
}

bool ffbsi_cand_tc_supported(const psvo_mlp& f, const psvo_mlp& g, int Dx, int Dy, long long rows) {
  static const bool disabled = getenv("PSVO_DISABLE_TC") != nullptr && getenv("PSVO_DISABLE_TC")[0] == '1';
  if (disabled) return false;
  if (f.n_hidden != 1 || g.n_hidden != 1) return false;
  const int H = g.width[0];
  if (f.width[0] != H || H < 128 || H > 256 || (H % 32) != 0) return false;
  if (Dx > 16 || Dy > 32) return false;
  return rows >= 128LL * kTcGroups * 32;
}

int launch_ffbsi_cand_tc(const CandTcArgs& a, cudaStream_t st) {
  PSVO_REQUIRE(a.tc_image != nullptr, "tensor-core FFBSi candidates: weight image missing");
  const TcPropSmem L = tc_prop_layout(a.H, 32);
  const size_t smem = L.total + 128;
  PSVO_REQUIRE(smem <= 227 * 1024, "tensor-core FFBSi candidates: %zu B of shared memory", smem);
  const long long tiles = (long long)a.B * (((long long)a.N * a.M + 127) / 128);
  long long grid = (tiles + kTcGroups - 1) / kTcGroups;
  if (grid > num_sms()) grid = num_sms();
  static bool cfg = false;
  if (!cfg) { PSVO_CUDA(cudaFuncSetAttribute(ffbsi_cand_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
  ProfScope ps_(PK_FFBSI, st);
  ffbsi_cand_tc_kernel<32><<<(int)grid, kTcThreads, smem, st>>>(a);
  PSVO_LAUNCH_CHECK();
  return 0;
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

size_t tc_image_bytes(const PropArgs& a) { return tc_prop_layout(a.g.width[0], 32).consts; }

int launch_tc_stage_image(const PropArgs& a, unsigned char* image, cudaStream_t st) {
  ProfScope ps_(PK_ROWS, st);
  tc_stage_image_kernel<32><<<1, 512, 0, st>>>(a, image);
  PSVO_LAUNCH_CHECK();
  return 0;
}

int launch_propagate_tc(const PropArgs& a, cudaStream_t st) {
  PSVO_REQUIRE(a.tc_image != nullptr, "tensor-core propagate: weight image missing");
  const int H = a.g.width[0];
  const int NPG = 32;      // one instantiation: the emission head is padded to 32 outputs
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const size_t smem = L.total + 128;
  PSVO_REQUIRE(smem <= 227 * 1024, "tensor-core propagate: %zu B of shared memory", smem);
  const int tiles = a.B * ((a.N + 127) / 128);
  int grid = (tiles + kTcGroups - 1) / kTcGroups;
  if (grid > num_sms()) grid = num_sms();
  ProfScope ps_(PK_PROP, st);
  {
    static bool cfg = false;      // (the trace build carries 8 B of static shared memory)
#ifdef PSVO_TC_TRACE
    const int max_dyn = 227 * 1024 - 1024;
#else
    const int max_dyn = 227 * 1024;
#endif
    if (!cfg) { PSVO_CUDA(cudaFuncSetAttribute(smc_propagate_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn)); cfg = true; }
    smc_propagate_tc_kernel<32><<<grid, kTcThreads, smem, st>>>(a);
  }
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo
