// smc_propagate_tc.cu -- forward-filter "propagate + weight" step with the per-particle MLPs on sm_100a tensor
// cores (tcgen05.mma, TMEM accumulators).  Same contract as smc_propagate_kernel (smc_forward.cu).
//
// Reference: SVO.SMC while_body (src/SMC/SVO.py:123-165) incl. sample_from_2_dist (:182-197,228-232); the two
// per-particle networks are MLP q1 (= f under use_bootstrap) and MLP g (src/transformation/MLP.py:48-54).
//
// Shape of the work (BASELINE config 5: Dx=16, Dy=32, H=256): per particle-step 41 kFLOP of tiny-K GEMMs
//   [128 x 16] x [16 x 256] -> relu -> [128 x 256] x [256 x 16|32].
//
// CTA = 4 compute groups of 128 threads + 4 MMA-issuing warps (one per group).  A group owns one 128-particle tile at a time
// (TMEM lane == row == thread) and walks the hidden layer of each network in stages of 16 units (one K-step of layer 2):
//     MMA warp g      : L1 of stage c (4 TS-form MMAs, N = 16: x_hi W_hi + x_hi W_lo + x_lo W_hi + bias) into one of FOUR layer-1
//                       accumulators, four stages ahead of the conversion; when stage c has been converted: L2 of stage c (3 MMAs,
//                       A operand = the converted accumulator itself) and L1 of stage c+4 into the very same columns
//     compute threads : wait L1(c) -> tcgen05.ld the accumulator row (prefetched one stage early) -> ReLU + fp16 hi/lo split in
//                       registers -> tcgen05.st the pair back IN PLACE -> mbarrier.arrive (published one stage late: no exposed
//                       wait::st).  The activations never touch shared memory; there is no block- or group-wide barrier in the loop.
// Every MMA is in TS form (A operand in tensor memory: the network's input row, the `ones` operand of the bias MMA, the
// converted activations), which costs N/2 + 1 pipe cycles down to N = 16 (profiles/microbench/r2_pipe_bench.cu); with the A
// operand in shared memory the floor is 32 + N/4 cycles (the 4 KB A tile at 128 B/clk).  All groups share the staged fp16
// weights (~96 KB at H = 256), copied by cp.async.bulk once per launch; DESIGN.md section 4a has the measurements.
// The same body (tc_prop_body) is the per-step kernel and the opt-in persistent whole-filter kernel (per-sequence flags in
// global memory instead of kernel boundaries, resampling inside the kernel: DESIGN.md section 4b).
// Arithmetic ("fp16x3"): x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 mantissa bits); a*w is evaluated as
// a_hi*w_hi + a_hi*w_lo + a_lo*w_hi with fp32 accumulation in TMEM; weights are pre-scaled by 2^8 (exact) so that
// their lo halves stay in fp16's normal range, and the scale is undone inside the bias FFMA.  Operands must stay
// below fp16's 65504 (a particle that large has diverged; it turns into NaN and flags its sequence via logZ_b).
#include <stdlib.h>
#include "mlp_tile.cuh"
#include "smc_kernels.h"
#include "tc_common.cuh"
#include "resample_group.cuh"

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

#ifdef PSVO_TC_PSTATS
// per-group cycle accounting of the persistent filter kernel (thread 0 of every compute group): 8 counters per group
//   0 wait for ready[b] | 1 tile body | 2 tile-end bookkeeping (barriers, atomics) | 3 lock wait | 4 resampling | 5 tiles | 6 resamplings
__device__ long long* g_tc_pstats = nullptr;
extern "C" __attribute__((visibility("default"))) int psvo_debug_tc_pstats(long long* buf) {
  cudaMemcpyToSymbol(g_tc_pstats, &buf, sizeof(buf));
  return 0;
}
#define PSTAT_DECL() long long ps_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pc_ = clock64()
#define PSTAT(k) do { const long long n_ = clock64(); ps_[k] += n_ - pc_; pc_ = n_; } while (0)
#define PSTAT_INC(k) do { ps_[k] += 1; } while (0)
#define PSTAT_DUMP(grp) do { if ((threadIdx.x & 127) == 0 && g_tc_pstats != nullptr) for (int k_ = 0; k_ < 8; ++k_) g_tc_pstats[(blockIdx.x * 4 + (grp)) * 8 + k_] = ps_[k_]; } while (0)
#else
#define PSTAT_DECL() do { } while (0)
#define PSTAT(k) do { } while (0)
#define PSTAT_INC(k) do { } while (0)
#define PSTAT_DUMP(grp) do { } while (0)
#endif

constexpr int kTcGroups = 4;
constexpr int kTcChunk = 16;           // hidden units per pipeline stage = one K-step of layer 2
constexpr int kTcBufs = 4;             // layer-1 accumulators per group
constexpr int kTcThreads = kTcGroups * 128 + kTcGroups * 32;
constexpr int kTcConsts = 160;                 // floats of one set of per-step constants (t = 0 | t > 0)
constexpr int kTcBars = 12;            // mbarriers per group
// exact power-of-two operand scales: W1 (and its bias) by 2^4, W2 by 2^8 -> their fp16 lo halves stay normal; the hidden
// activations travel in 16x units (|h| < 4094) and the layer-2 accumulator in 4096x units, undone in the output FFMA
constexpr float kW1Scale = 16.0f, kW2Scale = 256.0f, kWUnscale = 1.0f / (16.0f * 256.0f);

// Tensor-memory columns of one group (128 of the CTA's 512):
//   [0,32)    D2   layer-2 accumulator (q: 16 columns used, g: up to 32)
//   [32,48)   A1   the network's input row as the layer-1 A operand: fp16 hi (8 columns, two K-elements each) | lo (8 columns)
//   [48,56)   ones the A operand of the bias MMA: k = 0,1 -> 1, rest 0 (written once per launch)
//   [64,128)  D1   four layer-1 accumulators of 16 hidden units; the compute threads turn each IN PLACE into the layer-2
//                  A operand of one K-step: columns [0,8) fp16 hi (two hidden units per column), [8,16) fp16 lo
constexpr uint32_t kColD2 = 0, kColA1 = 32, kColOnes = 48, kColD1 = 64;
// mbarriers of one group: A1 staged (128 arrivals) | L1 of a stage complete, per buffer | stage converted (128), per buffer |
// L2 of a stage complete (only in the PSVO_TC_WAR_WAIT build) | last L2 of the network complete
enum { kBarA1 = 0, kBarD1 = 1, kBarFull = 5, kBarL2 = 9, kBarDone = 10, kBarImage = 11 /* group 0's slot: weight image landed */ };
// 1: wait for L2(c) to complete before L1(c+4) overwrites the accumulator it read its A operand from (conservative, slower)
#ifndef PSVO_TC_PUBLISH_FIRST
#define PSVO_TC_PUBLISH_FIRST 1
#endif
#ifndef PSVO_TC_WAR_WAIT
#define PSVO_TC_WAR_WAIT 0
#endif

struct TcPropSmem {
  uint32_t w1q_hi, w1q_lo, w1q_b, w2q_hi, w2q_lo;      // fp16 B operands of MLP q1 (f in the FFBSi candidate kernel)
  uint32_t w1g_hi, w1g_lo, w1g_b, w2g_hi, w2g_lo;      // ... of MLP g
  uint32_t b2q, b2g, consts;                           // fp32 output biases; per-launch constants (not part of the image)
  uint32_t rowc;                                       // per group, two tile parities: [mu2 (16) | obs (32) | pad] fp32
  uint32_t bars, slot, total;
};

__host__ __device__ inline TcPropSmem tc_prop_layout(int H, int NPG) {
  TcPropSmem s;
  uint32_t o = 0;
#define PSVO_TAKE(bytes) (o += (((bytes) + 127u) & ~127u), o - (((bytes) + 127u) & ~127u))
  s.w1q_hi = PSVO_TAKE(H * 16 * 2); s.w1q_lo = PSVO_TAKE(H * 16 * 2); s.w1q_b = PSVO_TAKE(H * 16 * 2);
  s.w2q_hi = PSVO_TAKE(16 * H * 2); s.w2q_lo = PSVO_TAKE(16 * H * 2);
  s.w1g_hi = PSVO_TAKE(H * 16 * 2); s.w1g_lo = PSVO_TAKE(H * 16 * 2); s.w1g_b = PSVO_TAKE(H * 16 * 2);
  s.w2g_hi = PSVO_TAKE(NPG * H * 2); s.w2g_lo = PSVO_TAKE(NPG * H * 2);
  s.b2q = PSVO_TAKE(16 * 4); s.b2g = PSVO_TAKE(32 * 4);
  s.consts = PSVO_TAKE(2 * kTcConsts * 4);
  s.rowc = PSVO_TAKE(kTcGroups * 2 * 64 * 4);
  s.bars = PSVO_TAKE(kTcBars * kTcGroups * 8); s.slot = PSVO_TAKE(16);
#undef PSVO_TAKE
  s.total = o;
  return s;
}

struct TcGroup {
  uint32_t bar;                    // shared-memory address of the group's mbarriers (kBar*, 8 bytes each)
  uint32_t par;                    // parities of the next completions this thread waits for: bits 0-3 = L1 into buffer 0-3, bit 4 = mDone
  uint32_t tb;                     // TMEM address of the group's column 0 incl. this warp's lane offset
};

// An opaque copy of a loop-invariant address.  Without it ptxas rematerialises the barrier / TMEM addresses from the kernel
// parameters at every use (a dependent chain of ~10 uniform-datapath instructions, ~100 cycles, in front of every
// mbarrier wait and arrive of the stage loop) instead of keeping them in a register.
__device__ __forceinline__ uint32_t tc_opaque(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n"
      :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mma_commit_u32(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait_bit(uint32_t bar, uint32_t& par, int bit) {
  mbar_wait_u32(bar, (par >> bit) & 1u);
  par ^= 1u << bit;
}

// ReLU + fp16 hi/lo split of one pair: hi = fp16(x) toward zero (ReLU inside the packing conversion), lo = fp16(x - hi) with the
// two subtractions in ONE packed fp32x2 instruction.  x < 0: hi = 0, x - hi < 0 -> lo = 0.  Five instructions per pair, two of
// them (the packing conversions) on the half-rate ALU pipe -- measured cheapest of the variants in profiles/microbench/r2_pipe_bench.cu.
__device__ __forceinline__ void split_pair_relu(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = tc::pack_relu_rz(x0, x1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  uint64_t a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x0), "f"(x1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(hf.x), "f"(hf.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  float l0, l1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(l0), "=f"(l1) : "l"(d));
  lo = tc::pack_relu_rn(l0, l1);
}

// One hidden-layer stage (16 hidden units) of the compute threads, software pipelined:
//   v  : the stage's accumulator row, its tcgen05.ld issued one stage earlier      -> wait::ld
//   vn : prefetch of stage c+1 (layer 1 runs three stages ahead, so its accumulator is normally complete)
//   ReLU + hi/lo split of v in registers; only then publish stage c-1 (its tcgen05.st has long completed: no exposed wait),
//   and store stage c IN PLACE as the layer-2 A operand of one K-step.  The activations never touch shared memory.
template <int BUF>
__device__ __forceinline__ void tc_stage(TcGroup& g, int c, int nst, uint32_t (&v)[16], uint32_t (&vn)[16]) {
  constexpr int NB = (BUF + 1) & (kTcBufs - 1), PB = (BUF + kTcBufs - 1) & (kTcBufs - 1);
  TC_TRACE(1);
  tc::tmem_ld16_wait(v);
  TC_TRACE(2);
#if PSVO_TC_PUBLISH_FIRST
  if (c > 0) {                                                             // publish stage c-1 BEFORE anything that can stall on the
    tc::tmem_st_wait();                                                    // MMA side: a late L1(c+1) must not hold back the arrival
    tc::fence_before_sync();                                               // that lets L1(c+3) be issued
    mbar_arrive_u32(g.bar + 8 * (kBarFull + PB));
  }
#endif
  if (c + 1 < nst) {
    mbar_wait_u32(g.bar + 8 * (kBarD1 + NB), (g.par >> NB) & 1u);          // L1 of stage c+1 complete
    g.par ^= 1u << NB;
    tc::fence_after_sync();
    tc::tmem_ld16_issue(g.tb + kColD1 + NB * kTcChunk, vn);
  }
  TC_TRACE(3);
#if !PSVO_TC_PUBLISH_FIRST
  if (c > 0) {                                                             // publish stage c-1: its tcgen05.st was issued a stage ago
    tc::tmem_st_wait();
    tc::fence_before_sync();
    mbar_arrive_u32(g.bar + 8 * (kBarFull + PB));
  }
#endif
  uint32_t r[16];
#pragma unroll
  for (int j = 0; j < 8; ++j)                                              // D1 = 16 (W1^T x + b1): the bias rode in the MMA
    split_pair_relu(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]), r[j], r[8 + j]);
  TC_TRACE(4);
  tc::tmem_st16(g.tb + kColD1 + BUF * kTcChunk, r);
  TC_TRACE(6);
}

// compute-group side of one network: x = this thread's input row.  On return the accumulator D2 is complete:
// D2 * kWUnscale = W2^T relu(W1^T x + b1).   nst = H / 16 stages (even).
__device__ __forceinline__ void tc_mlp_compute(TcGroup& g, const float (&x)[16], int nst) {
  {
    uint32_t r[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __half2 hh = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(x[2 * j] - hf.x, x[2 * j + 1] - hf.y);
      r[j] = *reinterpret_cast<const uint32_t*>(&hh);
      r[8 + j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    tc::tmem_st16(g.tb + kColA1, r);
    tc::tmem_st_wait();
  }
  tc::fence_before_sync();           // also orders this thread's earlier tcgen05.ld of the previous accumulators
  mbar_arrive_u32(g.bar + 8 * kBarA1);
  uint32_t va[16], vb[16];
  tc_wait_bit(g.bar + 8 * kBarD1, g.par, 0);                                   // L1 of stage 0 complete
  tc::fence_after_sync();
  tc::tmem_ld16_issue(g.tb + kColD1, va);
#pragma unroll 1
  for (int c = 0; c < nst; c += 4) {                                       // every network starts in buffer 0: static buffer indices
    tc_stage<0>(g, c, nst, va, vb);
    tc_stage<1>(g, c + 1, nst, vb, va);
    if (c + 2 < nst) {
      tc_stage<2>(g, c + 2, nst, va, vb);
      tc_stage<3>(g, c + 3, nst, vb, va);
    }
  }
  tc::tmem_st_wait();
  tc::fence_before_sync();
  mbar_arrive_u32(g.bar + 8 * (kBarFull + ((nst - 1) & (kTcBufs - 1))));         // last stage in place
  tc_wait_bit(g.bar + 8 * kBarDone, g.par, 4);                                 // L2(last) complete
  tc::fence_after_sync();
}

// ---- MMA-issuing warp ----------------------------------------------------------------------------------------------
// Everything the issuing thread touches is derived from warp-uniform values (shared-memory offsets of the layout, the loop
// counter) so that ptxas keeps descriptors and addresses in uniform registers: one event (wait -> 3 + 4 MMAs -> 2 commits)
// must cost far less than the conversion of one stage, or layer 1 stops running ahead of the compute threads.
struct TcNetAddr { uint32_t w1h, w1l, w1b, w2h, w2l; };     // shared-memory byte addresses of the network's B operands

// L1 of stage c -> D1 buffer `tD1` (4 TS-form MMAs, N = 16: x_hi W_hi + x_hi W_lo + x_lo W_hi + bias)
__device__ __forceinline__ void tc_issue_l1(const TcNetAddr& na, uint32_t lbo1, uint32_t tA1, uint32_t tD1, int c, uint32_t bar_d1) {
  constexpr uint32_t idesc1 = tc::make_idesc_f16(128, kTcChunk);
  const uint32_t off = (uint32_t)c * (kTcChunk * 16);
  const uint64_t w1h = tc::make_desc(na.w1h + off, lbo1, 128), w1l = tc::make_desc(na.w1l + off, lbo1, 128);
  const uint64_t w1b = tc::make_desc(na.w1b + off, lbo1, 128);
  tc::mma_f16_ts(tD1, tA1, w1h, idesc1, false);
  tc::mma_f16_ts(tD1, tA1, w1l, idesc1, true);
  tc::mma_f16_ts(tD1, tA1 + 8, w1h, idesc1, true);
  tc::mma_f16_ts(tD1, tA1 + 16, w1b, idesc1, true);                       // + 16 b1 (hi + lo); A = the ones columns
  mma_commit_u32(bar_d1);
}
// L2 of stage c = K-step c of layer 2, A operand in TMEM (the converted D1 buffer):  a_hi w_hi + a_hi w_lo + a_lo w_hi  into D2.
// A TS-form MMA occupies the tensor pipe for N/2 cycles (profiles/microbench/r2_pipe_bench.cu), so three N = NP instructions cost
// what the N-stacked pair of round 1 did and the accumulator needs no add of two halves.
template <int NP>
__device__ __forceinline__ void tc_issue_l2(const TcNetAddr& na, uint32_t tD1, uint32_t tD2, int c) {
  constexpr uint32_t idesc2 = tc::make_idesc_f16(128, NP);
  const uint32_t off = (uint32_t)(2 * c) * (NP * 16);                     // K-groups 2c, 2c+1
  const uint64_t w2h = tc::make_desc(na.w2h + off, NP * 16, 128), w2l = tc::make_desc(na.w2l + off, NP * 16, 128);
  tc::mma_f16_ts(tD2, tD1, w2h, idesc2, c > 0);
  tc::mma_f16_ts(tD2, tD1, w2l, idesc2, true);
  tc::mma_f16_ts(tD2, tD1 + 8, w2h, idesc2, true);
}

// one network of one tile:  wait A1 -> L1(0..3) ; per stage c: wait "stage c converted" -> L2(c) -> L1(c+4) into the very buffer
// L2(c) reads its A operand from.  Layer 1 thus runs four stages ahead of the conversion, which is what it takes to cover the
// round trip  arrive -> this warp wakes -> issue -> tensor pipe -> commit -> compute threads wake  (~500-600 cycles against
// ~300 per 16-unit stage), and the compute threads' prefetch finds its accumulator complete.
// Write-after-read on the D1 buffer: L1(c+4) is issued by this thread right after L2(c) and the tensor pipe executes one
// thread's tcgen05.mma in issue order (L2(c) has fetched its A operand before L1(c+4) can write the columns); waiting for
// L2(c)'s completion instead would put ~300 cycles into every one of the 64 events of a tile (PSVO_TC_WAR_WAIT=1 builds that
// conservative variant; it was the slower one by 25 %).
// par: bit 0 = A1, bits 1-4 = full[0..3], bit 5 = L2
template <int NP, bool WAIT_A1 = true>
__device__ __forceinline__ void tc_mma_net(const TcNetAddr& na, uint32_t lbo1, uint32_t tb, uint32_t bars, int nst, uint32_t& par) {
  TC_TRACE(10);
  if (WAIT_A1) {
    mbar_wait_u32(bars + 8 * kBarA1, par & 1u); par ^= 1u;
    tc::fence_after_sync();
  }
  TC_TRACE(11);
  if (tc::elect_one()) {
#pragma unroll
    for (int c = 0; c < kTcBufs; ++c)
      if (c < nst) tc_issue_l1(na, lbo1, tb + kColA1, tb + kColD1 + c * kTcChunk, c, bars + 8 * (kBarD1 + c));
  }
  __syncwarp();
#pragma unroll 1
  for (int c0 = 0; c0 < nst; c0 += kTcBufs) {
#pragma unroll
    for (int buf = 0; buf < kTcBufs; ++buf) {                              // static buffer index: addresses fold into immediates
      const int c = c0 + buf;
      if (c < nst) {
        mbar_wait_u32(bars + 8 * (kBarFull + buf), (par >> (1 + buf)) & 1u); par ^= 2u << buf;
        tc::fence_after_sync();
        TC_TRACE(12);
        const bool more = c + kTcBufs < nst;
        if (tc::elect_one()) {
          tc_issue_l2<NP>(na, tb + kColD1 + buf * kTcChunk, tb + kColD2, c);
          if (more) {
#if PSVO_TC_WAR_WAIT
            mma_commit_u32(bars + 8 * kBarL2);
            mbar_wait_u32(bars + 8 * kBarL2, (par >> 5) & 1u);             // L2(c) done -> its buffer reusable
            tc::fence_after_sync();
#endif
            tc_issue_l1(na, lbo1, tb + kColA1, tb + kColD1 + buf * kTcChunk, c + kTcBufs, bars + 8 * (kBarD1 + buf));
          } else if (c + 1 == nst) {
            mma_commit_u32(bars + 8 * kBarDone);
          }
        }
#if PSVO_TC_WAR_WAIT
        if (more) par ^= 32u;
#endif
        TC_TRACE(13);
        __syncwarp();
      }
    }
  }
}

// MMA-issuing warp of group `grp`: every tile of the group, networks first_net..1.  items = tiles of the launch (static stride);
// the persistent whole-filter kernel hands tiles out dynamically and tells this warp per tile what to run (see below)
template <int NPG, bool PERSIST = false>
__device__ __forceinline__ void tc_mma_warp(uint32_t tmem, uint32_t sm_base, const TcPropSmem& L, int H, int grp, int items, int nst, int first_net,
                                            int work = 0, bool skip_g = false) {
  const uint32_t bars = tc_opaque(sm_base + L.bars + 8u * kTcBars * grp);
  const uint32_t tb = tc_opaque(tmem + grp * 128);
  const uint32_t lbo1 = tc_opaque((uint32_t)H * 16u);
  TcNetAddr aq, ag;
  aq.w1h = tc_opaque(sm_base + L.w1q_hi); aq.w1l = tc_opaque(sm_base + L.w1q_lo); aq.w1b = tc_opaque(sm_base + L.w1q_b);
  aq.w2h = tc_opaque(sm_base + L.w2q_hi); aq.w2l = tc_opaque(sm_base + L.w2q_lo);
  ag.w1h = tc_opaque(sm_base + L.w1g_hi); ag.w1l = tc_opaque(sm_base + L.w1g_lo); ag.w1b = tc_opaque(sm_base + L.w1g_b);
  ag.w2h = tc_opaque(sm_base + L.w2g_hi); ag.w2l = tc_opaque(sm_base + L.w2g_lo);
  uint32_t par = 0;
  if constexpr (PERSIST) {
    // tiles are handed out dynamically: the compute group leaves a command word next to its first operand (published by the
    // A1 barrier's release / acquire): 0 = both networks, 1 = emission network only (t = 0), 2 = no more tiles
    const uint32_t cmd_addr = tc_opaque(sm_base + (uint32_t)work);         // `work` carries the shared-memory offset of the group's word
    for (;;) {
      mbar_wait_u32(bars + 8 * kBarA1, par & 1u); par ^= 1u;
      tc::fence_after_sync();
      uint32_t cmd;
      asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(cmd) : "r"(cmd_addr) : "memory");
      if (cmd == 2u) break;
      if (cmd == 0u) {
        tc_mma_net<16, false>(aq, lbo1, tb, bars, nst, par);
        tc_mma_net<NPG>(ag, lbo1, tb, bars, nst, par);
      } else {
        tc_mma_net<NPG, false>(ag, lbo1, tb, bars, nst, par);
      }
    }
  } else {
    for (int wi = blockIdx.x * kTcGroups + grp; wi < items; wi += gridDim.x * kTcGroups) {
      if (first_net == 0) tc_mma_net<16>(aq, lbo1, tb, bars, nst, par);
      if (!skip_g) tc_mma_net<NPG>(ag, lbo1, tb, bars, nst, par);
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
  tc::stage_weight_f16x2(a.q1.W[1], H, Dx, H, 16, kW2Scale, img + L.w2q_hi, img + L.w2q_lo);
  tc::stage_weight_f16x2(a.g.W[0], Dx, H, 16, H, kW1Scale, img + L.w1g_hi, img + L.w1g_lo);
  tc::stage_weight_f16x2(a.g.W[1], H, Dy, H, NPG, kW2Scale, img + L.w2g_hi, img + L.w2g_lo);
  // layer-1 biases as B operands: row n of K-group 0 = [fp16(16 b), fp16(16 b - hi), 0 ...]; K-group 1 stays zero
  for (int n = tid; n < H; n += blockDim.x) {
    const float bq = __ldg(a.q1.b[0] + n) * kW1Scale, bg = __ldg(a.g.b[0] + n) * kW1Scale;
    const __half qh = __float2half_rn(bq), gh = __float2half_rn(bg);
    const __half2 q2 = __halves2half2(qh, __float2half_rn(bq - __half2float(qh)));
    const __half2 g2 = __halves2half2(gh, __float2half_rn(bg - __half2float(gh)));
    *reinterpret_cast<__half2*>(img + L.w1q_b + (size_t)n * 16) = q2;
    *reinterpret_cast<__half2*>(img + L.w1g_b + (size_t)n * 16) = g2;
  }
  float* b2q = reinterpret_cast<float*>(img + L.b2q);
  float* b2g = reinterpret_cast<float*>(img + L.b2g);
  for (int i = tid; i < 16; i += blockDim.x) b2q[i] = i < Dx ? __ldg(a.q1.b[1] + i) : 0.0f;
  for (int i = tid; i < 32; i += blockDim.x) b2g[i] = i < Dy ? __ldg(a.g.b[1] + i) : 0.0f;
}

// ---- common prologue of the three kernels that run the pipeline -------------------------------------------------------
struct TcCtx { uint32_t tmem; uint64_t* bars; };

// part 1 (before the kernel's own constant staging): weight image -> shared memory (one coalesced copy of what
// tc_stage_image_kernel prepared once per pass), mbarrier init, TMEM allocation
__device__ __forceinline__ void tc_setup_begin(unsigned char* sm, const TcPropSmem& L, const unsigned char* __restrict__ image, int tid, int warp) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  if (tid == 0) {                                                          // one thread streams the image with bulk copies (no
    uint64_t* ib = bars + kBarImage;                                       // per-thread load/store instructions), counted on an mbarrier
    tc::mbar_init(ib, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tc::mbar_arrive_expect_tx(ib, L.consts);
    for (uint32_t off = 0; off < L.consts; off += 32768u) {
      const uint32_t n = L.consts - off < 32768u ? L.consts - off : 32768u;
      tc::bulk_g2s(sm + off, image + off, n, ib);
    }
  }
  if (tid < kTcBars * kTcGroups && tid != kBarImage) {
    const int k = tid % kTcBars;
    tc::mbar_init(bars + tid, (k == kBarA1 || (k >= kBarFull && k < kBarFull + kTcBufs)) ? 128 : 1);
  }
  if (warp == 0) tc::tmem_alloc(reinterpret_cast<uint32_t*>(sm + L.slot), 512);
}
// part 2: publish (generic -> async proxy for the staged weights, barrier init, TMEM address), build the operand descriptors
template <int NPG>
__device__ __forceinline__ void tc_setup_finish(unsigned char* sm, const TcPropSmem& L, int H, TcCtx& cx) {
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  cx.tmem = *reinterpret_cast<uint32_t*>(sm + L.slot);
  cx.bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  tc::mbar_wait(cx.bars + kBarImage, 0);                                   // the weight image has landed
}
// a compute thread's view of its group; also writes the thread's row of the bias MMA's "ones" operand (published by the
// thread's first arrival on the A1 barrier)
__device__ __forceinline__ void tc_group_init(TcGroup& g, const TcCtx& cx, int grp, int warp) {
  g.bar = tc_opaque(tc::smem_u32(cx.bars + kTcBars * grp));
  g.par = 0;
  g.tb = tc_opaque(cx.tmem + grp * 128 + ((uint32_t)((warp & 3) * 32) << 16));
  const uint32_t ones[8] = {0x3C003C00u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};     // half2(1, 1): k = 0 multiplies b_hi, k = 1 b_lo
  tc::tmem_st8(g.tb + kColOnes, ones);
  tc::tmem_st_wait();
}

// One body, two kernels.  PERSIST = false: one time step per launch (a.t), 640 threads.  PERSIST = true: the WHOLE filter in one
// launch -- every group walks the tiles of all T steps in one global order (item = t * work + tile; a group takes every
// (gridDim * 4)-th item, so the 7-vs-6 tiles-per-group imbalance of a single step averages out), weights / TMEM / barriers are set
// up once, and the step boundary is replaced by per-sequence flags in global memory:
//   cnt[b]   += 1 (acq_rel) when a group has stored a tile of sequence b.  The group whose increment completes step t of the
//              sequence (cnt[b] = tiles_per_seq * (t + 1)) resamples (b, t) right away (resample_group.cuh: 128 threads,
//              ~16 k warp-instructions, a quarter of the 1024-thread kernel's) in the CTA's one CDF region (shared-memory
//              lock), then ready[b] = t + 1 (release)
//   a tile of (b, t >= 1) starts with an acquire load of ready[b] >= t
// so the resampling of a sequence overlaps the propagation of all the others instead of owning the machine between two launches.
// (Every register of the SM belongs to the 16 compute + 4 issuing warps -- 4 x 112 + 32 = 480 of the 512 per scheduler that a
// launch can allocate -- so a dedicated resampling warp would have to take its registers from the compute warps.)
// No grid-wide barrier.  Progress: an item depends only on items with smaller numbers (directly, or through the resampling its
// owner group runs after it), every group handles its items in increasing order, the lock is only held across a resampling,
// which waits for nothing, and all CTAs are co-resident (cooperative launch).
// A bounded spin (PSVO watchdog) turns a protocol bug into NaN objectives instead of a hung GPU.
struct TcWait {
  int* err; bool dead;
  __device__ __forceinline__ void until(const int* flag, int target) {
    if (dead) return;
    unsigned spins = 0;
    while (ld_acquire_gpu(flag) < target) {
      __nanosleep(64);
      if ((++spins & 1023u) == 0u && (spins > (1u << 23) || *reinterpret_cast<volatile int*>(err) != 0)) {
        *reinterpret_cast<volatile int*>(err) = 1; dead = true; return;
      }
    }
  }
};

template <int NPG, bool PERSIST>
__device__ __forceinline__ void tc_prop_body(const PropArgs& a, const TcPersistArgs* pp, unsigned char* sm) {
  const int Dx = a.Dx, Dy = a.Dy, N = a.N, H = a.g.width[0];
  const bool two_q = a.flags & PSVO_F_TWO_Q, poisson = a.flags & PSVO_F_POISSON;
  const bool t0_launch = a.t == 0;
  const TcPropSmem L = tc_prop_layout(H, NPG);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform (kept in a uniform register)
  const int nchunk = H / kTcChunk;

  float* sb2q = reinterpret_cast<float*>(sm + L.b2q);
  float* sb2g = reinterpret_cast<float*>(sm + L.b2g);
  // consts: s1[16] | sp[16] | inv1[16] | inv2[16] | inv_sp[16] | inv_s1[16] | inv_sg[32] | ld_p ld_f ld_g
  float* C = reinterpret_cast<float*>(sm + L.consts);
  tc_setup_begin(sm, L, a.tc_image, tid, warp);
  // set 0: the launch's step (PERSIST: t = 0), set 1 (PERSIST only): t > 0
  if (tid < 16 || (PERSIST && tid >= 64 && tid < 80)) {
    const int d = tid & 15, set = tid >> 6;
    const bool first = PERSIST ? set == 0 : t0_launch;
    float* Cs = C + set * kTcConsts;
    const float s1 = a.sig[(first ? SIG_Q0 : SIG_Q1) * 32 + d], s2 = a.sig[SIG_Q2 * 32 + d];
    const float i1 = 1.0f / s1, i2 = 1.0f / s2;
    const float spv = two_q ? 1.0f / (i1 + i2) : s1;
    Cs[d] = s1; Cs[16 + d] = spv; Cs[32 + d] = i1; Cs[48 + d] = i2; Cs[64 + d] = 1.0f / spv; Cs[80 + d] = i1;
  }
  if ((tid >= 32 && tid < 64) || (PERSIST && tid >= 96 && tid < 128)) {
    const int d = tid & 31, set = tid >> 6;
    C[set * kTcConsts + 96 + d] = d < Dy ? 1.0f / a.sig[SIG_G * 32 + d] : 0.0f;   // 0 beyond Dy: padded columns drop out
  }
  TC_TRACE_INIT();
  TcCtx tcx;
  tc_setup_finish<NPG>(sm, L, H, tcx);
  if (tid == 0 || (PERSIST && tid == 64)) {
    float* Cs = C + (tid >> 6) * kTcConsts;
    float lp = 0.f, lf = 0.f, lg = 0.f;
    for (int d = 0; d < Dx; ++d) { lp += logf(Cs[16 + d]); lf += logf(Cs[d]); }
    if (!poisson) for (int d = 0; d < Dy; ++d) lg += logf(a.sig[SIG_G * 32 + d]);
    Cs[128] = lp; Cs[129] = lf; Cs[130] = lg;
  }
  __syncthreads();
  const int tiles_per_seq = (N + 127) / 128;
  const int work = a.B * tiles_per_seq;
  const int items = PERSIST ? work * a.T : work;                         // < 2^31 (checked by the launcher)

  // programmatic dependent launch: everything above (weight image, barriers, TMEM) ran while the previous kernel (the resampling
  // of step t-1) was still draining; the next kernel may start its own prologue from here on, and the first read of anything the
  // previous kernel wrote (ancestors) comes after griddep_wait()
  if (!PERSIST) griddep_launch_dependents();
  if (warp >= kTcGroups * 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");                // 128 x (96 - 32) registers released = exactly the 512 x (112 - 96) the compute warps take
    tc_mma_warp<NPG, PERSIST>(tcx.tmem, tc::smem_u32(sm), L, H, warp - kTcGroups * 4, items, nchunk, t0_launch ? 1 : 0,
                              PERSIST ? (int)(pp->rs_smem_off + 4u * (12 + warp - kTcGroups * 4)) : 0);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    if (!PERSIST) griddep_wait();
    // =================================== compute groups ===================================
    const int grp = warp >> 2;
    const float* inv_sg = C + 96;                                          // the emission constants do not depend on the step
    const float ld_g = C[130];
    const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;
    const float neg_logN = -logf((float)N);
    TcWait wt{PERSIST ? pp->err : nullptr, false};
    TcGroup g;
    const int gt = tid & 127;
    tc_group_init(g, tcx, grp, warp);

    uint32_t tile_it = 0;
    // PERSIST: region header words: 0 lock | 1..4 "this group resamples" | 8..11 the group's next item | 12..15 command word of
    // the group's MMA warp.  Items are handed out by a global counter, and a group only takes its next one when it is free to
    // start it: a tile held across a resampling would finish last in ITS sequence, make this group resample that one too, and so
    // on (measured: chains of tile + resampling per group, every sequence they touch late by two resamplings)
    int* rg = PERSIST ? reinterpret_cast<int*>(sm + pp->rs_smem_off) : nullptr;
    int item = blockIdx.x * kTcGroups + grp;
    if (PERSIST) {
      if (gt == 0) rg[8 + grp] = atomicAdd(pp->next, 1);
      rg_bar(1 + grp);
      item = rg[8 + grp];
    }
    PSTAT_DECL();
    for (; item < items; item += PERSIST ? 0 : gridDim.x * kTcGroups) {
      const int t = PERSIST ? item / work : a.t;
      const int wi = PERSIST ? item - t * work : item;
      const bool t0 = t == 0;
      if (PERSIST && gt == 0) rg[12 + grp] = t0 ? 1 : 0;
      const float* Cs = C + ((PERSIST && !t0) ? kTcConsts : 0);
      const float* sp = Cs + 16; const float* inv1 = Cs + 32; const float* inv2 = Cs + 48;
      const float* inv_sp = Cs + 64; const float* inv_s1 = Cs + 80;
      const float ld_p = Cs[128], ld_f = Cs[129];
      const int b = wi / tiles_per_seq;
      const int n = (wi - b * tiles_per_seq) * 128 + gt;
      const bool live = n < N;
      const size_t bt = (size_t)b * a.T + t;
      PSTAT(2);
      if (PERSIST && !t0) wt.until(pp->ready + b, t);                      // (b, t-1) resampled: its ancestors are visible (acquire)
      PSTAT(0);
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
        xsrc = a.anc_prev != nullptr ? a.Xs + ((bt - 1) * N + (PERSIST ? __ldcg(a.anc_prev + prow) : __ldg(a.anc_prev + prow))) * Dx : a.X_anc + prow * Dx;
      }
      // noise for this particle (the Philox work also covers the latency of the ancestor index load above)
      float e[16];
      if (a.eps != nullptr) {
        const float* src = a.eps + row * Dx;
#pragma unroll
        for (int d = 0; d < 16; ++d) e[d] = d < Dx ? src[d] : 0.0f;
      } else {
        const uint64_t gidx = ((uint64_t)(a.seq_offset + b) * a.T + t) * N + (live ? n : 0);
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
        const uint64_t gidx = ((uint64_t)(a.seq_offset + b) * a.T + t) * N + (live ? n : 0);
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
        tc::tmem_ld16(g.tb + kColD2, mu1);
#pragma unroll
        for (int d = 0; d < 16; ++d) mu1[d] = fmaf(mu1[d], kWUnscale, sb2q[d]);
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
        float mug[16];
        tc::tmem_ld16(g.tb + kColD2 + c, mug);
        if (!poisson) {                                                    // branch-free Gaussian emission; columns beyond Dy: y = b2 = 0
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int d = c + i;
            const float mu = fmaf(mug[i], kWUnscale, sb2g[d]);
            const float z = (rc[16 + d] - mu) * inv_sg[d];
            gsum = fmaf(z, z, gsum);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int d = c + i;
            if (d < Dy) {
              const float mu = fmaf(mug[i], kWUnscale, sb2g[d]);
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
      if (PERSIST) {
        // tile stored: the group's barrier orders the 128 threads' stores before thread 0's acq_rel increment (cumulative at
        // gpu scope).  Whoever completes the step of the sequence resamples it -- the whole group, ~16 k warp-instructions
        // (t = T-1 is left to the resampling launch that follows the kernel: it also gathers X_anc[T-1])
        PSTAT(1); PSTAT_INC(5);
        rg_bar(1 + grp);
        if (gt == 0) {
          const int last = t + 1 < a.T && atom_add_acq_rel_gpu(pp->cnt + b, 1) + 1 == tiles_per_seq * (t + 1);
          rg[1 + grp] = last;
          if (!last) rg[8 + grp] = atomicAdd(pp->next, 1);
        }
        rg_bar(1 + grp);
        if (rg[1 + grp]) {
          PSTAT(2);
          if (gt == 0) while (atomicCAS(rg, 0, 1) != 0) __nanosleep(200);
          rg_bar(1 + grp);
          PSTAT(3);
          resample_group(pp->ra, b, t, sm + pp->rs_smem_off, gt, 1 + grp);
          rg_bar(1 + grp);                                                 // the group's ancestor stores happen before thread 0's release
          PSTAT(4); PSTAT_INC(6);
          if (gt == 0) {
            st_release_gpu(pp->ready + b, t + 1);
            __threadfence_block();
            atomicExch(rg, 0);
            rg[8 + grp] = atomicAdd(pp->next, 1);
          }
          rg_bar(1 + grp);
        }
        item = rg[8 + grp];
      }
    }
    PSTAT_DUMP(grp);
    if (PERSIST) {                                                         // no more tiles: release the group's MMA warp
      if (gt == 0) rg[12 + grp] = 2;
      mbar_arrive_u32(g.bar + 8 * kBarA1);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tcx.tmem, 512);
}

template <int NPG>
__global__ void __launch_bounds__(kTcThreads, 1) smc_propagate_tc_kernel(PropArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  tc_prop_body<NPG, false>(a, nullptr, sm);
}
template <int NPG>
__global__ void __launch_bounds__(kTcThreads, 1) smc_filter_tc_persistent_kernel(PropArgs a, TcPersistArgs p) {
  extern __shared__ __align__(128) unsigned char sm[];
  if (threadIdx.x < 16) reinterpret_cast<int*>(sm + p.rs_smem_off)[threadIdx.x] = 0;  // the CDF region's lock and flags (published by the prologue's barriers)
  tc_prop_body<NPG, true>(a, &p, sm);
  if (threadIdx.x == 0 && *reinterpret_cast<volatile int*>(p.err) != 0)          // watchdog fired somewhere: poison the objective
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) p.ra.logZ_b[b] = __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------------------------------
// FFBSi candidates: MLP f and MLP g of every candidate x~[m,i] on tensor cores  (PSVO.py:125,130 -- f.log_prob(bw_X_t,
// bw_X_tp1) and g.log_prob(bw_X_t, obs[:, t]) on M*N*B rows).  Same pipeline as smc_propagate_tc_kernel (the weight
// image holds f in the "q" slot); the candidate is recomputed from eps_bw / Philox exactly like ffbsi_step_kernel does.
// Outputs f_lp and g_lp per candidate, row index (b, i, m).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fwd2_load_row(const float* __restrict__ src, int Dx, bool live, float (&x)[16]);

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
  tc_setup_begin(sm, L, a.tc_image, tid, warp);
  if (tid < 16) {
    const int d = tid;
    const float s1 = a.sig[(last ? SIG_BQI : SIG_Q1I) * 32 + d], s2 = a.sig[SIG_BQ2 * 32 + d];
    C[16 + d] = 1.0f / s1; C[32 + d] = 1.0f / s2;
    C[d] = last ? s1 : 1.0f / (1.0f / s1 + 1.0f / s2);
    C[100 + d] = 1.0f / C[d];
    C[48 + d] = 1.0f / a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d];
  }
  if (tid >= 32 && tid < 64) C[64 + tid - 32] = tid - 32 < Dy ? 1.0f / a.sig[SIG_G * 32 + tid - 32] : 0.0f;   // 0 beyond Dy: padded columns drop out
  TcCtx tcx;
  tc_setup_finish<NPG>(sm, L, H, tcx);
  if (tid == 0) {
    float lf = 0.f, lg = 0.f, lq = 0.f;
    for (int d = 0; d < Dx; ++d) { lf += logf(a.sig[(boot ? SIG_Q1 : SIG_F) * 32 + d]); lq += logf(C[d]); }
    if (!poisson) for (int d = 0; d < Dy; ++d) lg += logf(a.sig[SIG_G * 32 + d]);
    C[96] = lf; C[97] = lg; C[98] = lq;
  }
  __syncthreads();
  const long long ncand = (long long)N * M;
  const int tiles_per_seq = (int)((ncand + 127) / 128);
  const int work = a.B * tiles_per_seq;

  if (warp >= kTcGroups * 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
    tc_mma_warp<NPG>(tcx.tmem, tc::smem_u32(sm), L, H, warp - kTcGroups * 4, work, nchunk, last ? 1 : 0);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    const int grp = warp >> 2;
    const float* sp = C; const float* inv1 = C + 16; const float* inv2 = C + 32; const float* inv_sf = C + 48;
    const float* inv_sg = C + 64; const float* inv_sp = C + 100;
    const float ld_f = C[96], ld_g = C[97], ld_q = C[98];
    const float cx = 0.5f * (float)Dx * kLog2Pi, cy = 0.5f * (float)Dy * kLog2Pi;
    TcGroup g;
    const int gt = tid & 127;
    tc_group_init(g, tcx, grp, warp);

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
      float qs = 0.f;
      float m1v[16];                                                       // the trajectory's first proposal mean: one 128-bit row read
      if (!last) tc_fwd2_load_row(a.mu1i + ((size_t)b * N + i) * Dx, Dx, true, m1v);
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
              const float m1 = m1v[d], m2 = rc[d];
              mean = sp[d] * (inv1[d] * m1 + inv2[d] * m2);
            }
            xv = mean + sp[d] * e[k];
            const float z = (xv - mean) * inv_sp[d];                          // the proposal density of the candidate
            qs = fmaf(z, z, qs);
          }
          x[d] = xv;
        }
      }

      // ---- f(x_{t+1} | x~)
      float f_lp = 0.0f;
      if (!last) {
        tc_mlp_compute(g, x, nchunk);
        float muf[16];
        tc::tmem_ld16(g.tb + kColD2, muf);
        float xn[16];
        tc_fwd2_load_row(a.X_next + ((bt + 1) * N + i) * Dx, Dx, true, xn);
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          if (d < Dx) {
            const float mu = fmaf(muf[d], kWUnscale, sb2q[d]);
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
        float mug[16];
        tc::tmem_ld16(g.tb + kColD2 + c0, mug);
        if (!poisson) {                                                    // branch-free Gaussian emission; columns beyond Dy: y = b2 = 0
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int d = c0 + k;
            const float mu = fmaf(mug[k], kWUnscale, sb2g[d]);
            const float z = (rc[16 + d] - mu) * inv_sg[d];
            gsum = fmaf(z, z, gsum);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int d = c0 + k;
            if (d < Dy) {
              const float mu = fmaf(mug[k], kWUnscale, sb2g[d]);
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
        if (a.candQ != nullptr) a.candQ[(size_t)b * ncand + c] = -0.5f * qs - ld_q - cx;
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tcx.tmem, 512);
}

// ------------------------------------------------------------------------------------------------------
// Plain forward of the two per-particle networks for the reverse pass (smc_backward.cu recomputes the network outputs
// instead of storing them):  mu_q = MLP q1(xq rows), mu_g = MLP g(xg rows), same pipeline as smc_propagate_tc_kernel.
// ------------------------------------------------------------------------------------------------------
// one particle row in / 16 output columns out: 128-bit accesses when the width allows (rows are 16-byte aligned then)
__device__ __forceinline__ void tc_fwd2_load_row(const float* __restrict__ src, int Dx, bool live, float (&x)[16]) {
  if (Dx == 16) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float4 q = live ? reinterpret_cast<const float4*>(src)[v] : make_float4(0.f, 0.f, 0.f, 0.f);
      x[4 * v] = q.x; x[4 * v + 1] = q.y; x[4 * v + 2] = q.z; x[4 * v + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int d = 0; d < 16; ++d) x[d] = (live && d < Dx) ? src[d] : 0.0f;
  }
}
__device__ __forceinline__ void tc_fwd2_store_row(float* __restrict__ dst, int D, int c0, const float (&mu)[16], const float* __restrict__ bias) {
  if ((D & 3) == 0) {
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (c0 + 4 * v < D)
        reinterpret_cast<float4*>(dst + c0)[v] = make_float4(fmaf(mu[4 * v], kWUnscale, bias[c0 + 4 * v]), fmaf(mu[4 * v + 1], kWUnscale, bias[c0 + 4 * v + 1]),
                                                             fmaf(mu[4 * v + 2], kWUnscale, bias[c0 + 4 * v + 2]), fmaf(mu[4 * v + 3], kWUnscale, bias[c0 + 4 * v + 3]));
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) if (c0 + k < D) dst[c0 + k] = fmaf(mu[k], kWUnscale, bias[c0 + k]);
  }
}

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
  tc_setup_begin(sm, L, a.tc_image, tid, warp);
  TcCtx tcx;
  tc_setup_finish<NPG>(sm, L, H, tcx);
  const long long rows = (long long)a.B * N;
  const int work = (int)((rows + 127) / 128);

  if (warp >= kTcGroups * 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
    tc_mma_warp<NPG>(tcx.tmem, tc::smem_u32(sm), L, H, warp - kTcGroups * 4, work, nchunk, with_q ? 0 : 1, 0, a.xg == nullptr);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    const int grp = warp >> 2;
    TcGroup g;
    const int gt = tid & 127;
    tc_group_init(g, tcx, grp, warp);
    for (int wi = blockIdx.x * kTcGroups + grp; wi < work; wi += gridDim.x * kTcGroups) {
      const long long r = (long long)wi * 128 + gt;
      const bool live = r < rows;
      const int b = live ? (int)(r / N) : 0, n = live ? (int)(r - (long long)b * N) : 0;
      float x[16];
      if (with_q) {
        const float* src = a.xq + (long long)b * a.x_seq_stride + (long long)n * Dx;
        tc_fwd2_load_row(src, Dx, live, x);
        tc_mlp_compute(g, x, nchunk);
        float mu[16];
        tc::tmem_ld16(g.tb + kColD2, mu);
        if (live) tc_fwd2_store_row(a.muq + r * Dx, Dx, 0, mu, sb2q);
      }
      if (a.xg != nullptr) {
        const float* src = a.xg + (long long)b * a.x_seq_stride + (long long)n * Dx;
        tc_fwd2_load_row(src, Dx, live, x);
        tc_mlp_compute(g, x, nchunk);
#pragma unroll
        for (int c0 = 0; c0 < NPG; c0 += 16) {
          float mu[16];
          tc::tmem_ld16(g.tb + kColD2 + c0, mu);
          if (live) tc_fwd2_store_row(a.mug + r * Dy, Dy, c0, mu, sb2g);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tcx.tmem, 512);
}

int launch_tc_fwd2(const Fwd2TcArgs& a, cudaStream_t st) {
  PSVO_REQUIRE(a.tc_image != nullptr && ((a.xg != nullptr && a.mug != nullptr) || (a.xq != nullptr && a.muq != nullptr)), "tensor-core forward: null buffer");
  const TcPropSmem L = tc_prop_layout(a.H, 32);
  const size_t smem = L.total + 128;
  PSVO_REQUIRE(smem <= 227 * 1024, "tensor-core forward: %zu B of shared memory", smem);
  const long long tiles = ((long long)a.B * a.N + 127) / 128;
  long long grid = (tiles + kTcGroups - 1) / kTcGroups;
  if (grid > num_sms()) grid = num_sms();
  PSVO_DYN_SMEM(tc_fwd2_kernel<32>, 227 * 1024);
  ProfScope ps_(a.prof_kind, st);
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
  PSVO_DYN_SMEM(ffbsi_cand_tc_kernel<32>, 227 * 1024);
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
  // the tensor path pays off once the hidden layer dominates and the machine can be filled (PSVO_TC_MIN_ROWS: test hook that
  // lowers the row threshold so small reference-generated vectors can be run through the tensor-core kernels)
  long long min_rows = 128LL * kTcGroups * 32;
  if (const char* e = getenv("PSVO_TC_MIN_ROWS")) min_rows = atoll(e);
  if (H < 128 || (long long)a.B * a.N < min_rows) return false;
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
    // (the trace build carries 8 B of static shared memory)
#ifdef PSVO_TC_TRACE
    const int max_dyn = 227 * 1024 - 1024;
#else
    const int max_dyn = 227 * 1024;
#endif
    PSVO_DYN_SMEM(smc_propagate_tc_kernel<32>, max_dyn);
    PSVO_CUDA(launch_pdl(smc_propagate_tc_kernel<32>, dim3(grid), dim3(kTcThreads), smem, st, a.pdl != 0, a));
  }
  PSVO_LAUNCH_CHECK();
  return 0;
}

// ---- persistent whole-filter launch ------------------------------------------------------------------------------------
// opt-in (PSVO_TC_PERSIST=1): measured slower than per-step launches at the config-5 shard (the chain
// tile -> resampling -> next step's tile of a sequence is longer than the step period once the resampling competes with
// sixteen other warps for issue slots)
static bool tc_persist_env() {
  static const bool on = getenv("PSVO_TC_PERSIST") != nullptr && getenv("PSVO_TC_PERSIST")[0] == '1';
  return on;
}
size_t tc_filter_sync_bytes(int B) { return align_up(((size_t)2 * B + 32) * sizeof(int), 256); }

bool tc_filter_persistent_supported(const PropArgs& a, const ResampleArgs& ra) {
  if (!tc_persist_env() || a.tc_image == nullptr || !tc_propagate_supported(a)) return false;
  // resampling on (IWAE keeps its per-particle log Wbar in a buffer that is rewritten every step: per-step launches), at
  // least two steps, and the worker's CDF / guide table must fit next to the weight image
  if (!(a.flags & PSVO_F_RESAMPLE) || a.T < 2 || a.N < 128) return false;
  if ((long long)a.B * ((a.N + 127) / 128) * a.T >= (1LL << 31)) return false;
  (void)ra;
  const TcPropSmem L = tc_prop_layout(a.g.width[0], 32);
  return L.total + 256 + rg_smem_bytes(a.N) <= 227 * 1024;
}

int launch_filter_tc_persistent(PropArgs a, const ResampleArgs& ra, int* sync, cudaStream_t st) {
  PSVO_REQUIRE(tc_filter_persistent_supported(a, ra) && sync != nullptr, "persistent tensor-core filter: unsupported shape");
  const TcPropSmem L = tc_prop_layout(a.g.width[0], 32);
  TcPersistArgs p;
  p.ra = ra; p.ra.t = 0; p.ra.skip_gather = 1;
  p.cnt = sync; p.ready = sync + a.B; p.err = sync + 2 * a.B; p.next = sync + 2 * a.B + 1;
  p.rs_smem_off = (L.total + 255u) & ~127u;
  const size_t smem = (size_t)p.rs_smem_off + rg_smem_bytes(a.N);
  const int tiles = a.B * ((a.N + 127) / 128);
  int grid = (tiles + kTcGroups - 1) / kTcGroups;
  if (grid > num_sms()) grid = num_sms();
  a.t = 0; a.anc_prev = ra.anc; a.logWbar = nullptr;
  PSVO_CUDA(cudaMemsetAsync(sync, 0, tc_filter_sync_bytes(a.B), st));
  ProfScope ps_(PK_PROP, st);
  PSVO_DYN_SMEM(smc_filter_tc_persistent_kernel<32>, 227 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kTcThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;           // co-residency of all CTAs is part of the protocol: let the runtime check it
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  PSVO_CUDA(cudaLaunchKernelEx(&cfg, smc_filter_tc_persistent_kernel<32>, a, p));
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo
