// mlp_bwd_tc.cu -- reverse mode of a one-hidden-layer MLP_transformation over particle rows on sm_100a tensor cores.
//
// Reference semantics: TF autodiff through MLP_transformation.transform (src/transformation/MLP.py:48-54), i.e. what
// optimizer.minimize(-log_ZSMC) (src/trainer.py:115-118) derives for the per-particle networks q1 / f / g.  Given the
// rows x (rows, d_in <= 16) and the upstream gradient G = dL/dmu (rows, d_out <= 32):
//     pre = x W1 + b1,  h = relu(pre),  dh = G W2^T,  gpre = dh * [pre > 0]
//     gx  = gpre W1^T            dW2 = h^T G     db2 = sum_p G      dW1 = x^T gpre     db1 = sum_p gpre
// All five products run on tcgen05 with fp16x3 split operands and fp32 TMEM accumulators:
//     per 128-row tile and per chunk of 64 hidden units
//       L1  : D1[128 x 64] = x      * W1c          (K = 16)            row-wise, K-major operands
//       dh  : D3[128 x 64] = G      * W2c^T        (K = d_out)         row-wise
//       dx  : DX[128 x 16]+= gpre_c * W1c^T        (K = 64)            row-wise; B = the L1 operand W1 read MN-major
//       dW2 : AW2c[64 x NP]  += h_c^T    * G       (K = 128 rows)      MN-major operands (see below), M = 64
//       dW1 : AW1c[64 x 24]  += gpre_c^T * [x | 1] (K = 128 rows)      MN-major; the ones column yields db1
// The weight-gradient GEMMs contract over the ROWS.  Their operands are read straight from the row-wise buffers the
// epilogue writes anyway (thread = row stores 8 consecutive columns as one 16-byte chunk) through MN-major descriptors:
// element (mn, k) at (mn%8)*2 + (k%8)*16 + (k/8)*LBO + (mn/8)*SBO with LBO = 128 B, SBO = 2048 B (verified on B200 by
// profiles/microbench/umma_mnmajor_test.cu, which also gives the M = 64 accumulator lane map: row r -> lane (r/16)*32 + r%16).
// The weight-gradient accumulators stay in TMEM for all tiles of the (persistent) CTA and are flushed once.
//
// Scales (exact powers of two): W1 x16 (also read MN-major as the B operand of dx), W2 x256, G x sG with sG chosen from max|G| of the launch so that the
// fp16 lo halves stay normal; every accumulator is un-scaled in fp32 when it is read.
//
// CTA = 256 compute threads (TMEM lane == row; two threads per row split the columns of every chunk) + three MMA-issuing
// warps (row-wise products, dW2, dW1).  One CTA per SM (the accumulators take the
// TMEM): this first version overlaps the big weight-gradient MMAs of chunk c with the epilogue arithmetic of chunk c+1
// through the early issue of L1/dh and through double-buffered h / gpre operands.
#include <stdlib.h>
#include "smc_kernels.h"
#include "tc_common.cuh"

namespace psvo {

constexpr int kBtChunk = 64;
// stride between the 8-column groups of a row-operand buffer: 128 rows x 16 B plus 16 B of padding, so that the 8 (or 4, 3)
// column groups an MN-major MMA reads for one k land in different shared-memory banks (no-swizzle layout)
constexpr int kBtGrp = 2048 + 16;
constexpr int kBtBuf = 8 * kBtGrp;       // one chunk (64 columns) of a row operand
constexpr float kBtW1Scale = 16.0f, kBtW2Scale = 256.0f;

struct BtSmem {
  uint32_t w1_hi, w1_lo, w2_hi, w2_lo, b1s, image_end;
  uint32_t x_hi, x_lo, g_hi, g_lo, hc_hi, hc_lo, gp_hi, gp_lo, db2, bars, slot, total;
};
__host__ __device__ inline BtSmem bt_layout(int H, int NP) {
  BtSmem s; uint32_t o = 0;
#define PSVO_TAKE(bytes) (o += (((bytes) + 127u) & ~127u), o - (((bytes) + 127u) & ~127u))
  s.w1_hi = PSVO_TAKE(H * 32); s.w1_lo = PSVO_TAKE(H * 32);
  s.w2_hi = PSVO_TAKE(H * NP * 2); s.w2_lo = PSVO_TAKE(H * NP * 2);
  s.b1s = PSVO_TAKE(H * 4);
  s.image_end = o;
  s.x_hi = PSVO_TAKE(3 * kBtGrp); s.x_lo = PSVO_TAKE(3 * kBtGrp);
  s.g_hi = PSVO_TAKE((NP / 8) * kBtGrp); s.g_lo = PSVO_TAKE((NP / 8) * kBtGrp);
  s.hc_hi = PSVO_TAKE(2 * kBtBuf); s.hc_lo = PSVO_TAKE(2 * kBtBuf);     // double-buffered over chunks
  s.gp_hi = PSVO_TAKE(2 * kBtBuf); s.gp_lo = PSVO_TAKE(2 * kBtBuf);
  s.db2 = PSVO_TAKE(32 * 4);
  s.bars = PSVO_TAKE(8 * 8); s.slot = PSVO_TAKE(16);
#undef PSVO_TAKE
  s.total = o;
  return s;
}

// B operand [ROWS x K] (K-major, no swizzle) from a row-major matrix: B(r, k) = W[r*rs + k*cs] * scale, hi/lo split
__device__ __forceinline__ void bt_stage_b(const float* __restrict__ W, int rows, int kdim, int rs, int cs, int ROWS, int KP, float scale,
                                           unsigned char* hi, unsigned char* lo) {
  for (int i = threadIdx.x; i < ROWS * KP; i += blockDim.x) {
    const int r = i % ROWS, k = i / ROWS;
    const float w = (r < rows && k < kdim) ? __ldg(W + (size_t)r * rs + (size_t)k * cs) * scale : 0.0f;
    const __half h = __float2half_rn(w);
    const __half l = __float2half_rn(w - __half2float(h));
    const size_t off = (size_t)(k >> 3) * (ROWS * 16) + (size_t)r * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(hi + off) = h;
    *reinterpret_cast<__half*>(lo + off) = l;
  }
}

__global__ void __launch_bounds__(512) mlp_bwd_tc_stage_kernel(psvo_mlp net, int NP, unsigned char* __restrict__ img) {
  const int H = net.width[0], din = net.d_in, dout = net.d_out;
  const BtSmem L = bt_layout(H, NP);
  bt_stage_b(net.W[0], H, din, 1, H, H, 16, kBtW1Scale, img + L.w1_hi, img + L.w1_lo);         // (j, d) = W1[d][j]
  bt_stage_b(net.W[1], H, dout, dout, 1, H, NP, kBtW2Scale, img + L.w2_hi, img + L.w2_lo);     // (j, o) = W2[j][o]
  float* b1s = reinterpret_cast<float*>(img + L.b1s);
  for (int i = threadIdx.x; i < H; i += blockDim.x) b1s[i] = __ldg(net.b[0] + i) * kBtW1Scale;
}

// max |G| over the launch (positive floats order like their bit patterns)
__global__ void absmax_kernel(const float* __restrict__ g, long long n, unsigned int* __restrict__ out) {
  float m = 0.0f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(g[i]));
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out, __float_as_uint(m));
}

struct BtArgs {
  const unsigned char* image;
  const float* x; long long x_seq_stride;   // row (b, n) at x + b*x_seq_stride + n*d_in
  const float* g;                           // (rows, d_out) contiguous
  const unsigned int* gmax_bits;            // max |g| of the launch
  float* gx;                                // (rows, d_in): written, or scattered with atomicAdd through `scatter`
  const int* scatter; long long scatter_seq_stride;   // parent index of row (b, n) at scatter + b*stride + n, or NULL
  psvo_mlp_grad grads;
  int B, N, H, NP, d_in, d_out;
};

__device__ __forceinline__ void bt_mma3(uint32_t d, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t idesc, bool acc) {
  tc::mma_bf16(d, ah, bh, idesc, acc);
  tc::mma_bf16(d, ah, bl, idesc, true);
  tc::mma_bf16(d, al, bh, idesc, true);
}

__global__ void __launch_bounds__(352, 1) mlp_bwd_tc_kernel(BtArgs a) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int H = a.H, NP = a.NP, din = a.d_in, dout = a.d_out, N = a.N;
  const BtSmem L = bt_layout(H, NP);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int nch = H / kBtChunk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + L.bars);
  uint64_t* xg_full = bars;        // 128 arrivals : x / G operands of the tile staged
  uint64_t* d_ready = bars + 1;    // commit       : L1(c) + dh(c) complete
  uint64_t* d_read = bars + 2;     // 128 arrivals : D1 / D3 of chunk c are in registers
  uint64_t* a_ready = bars + 3;    // 128 arrivals : h / gpre operands of chunk c staged
  uint64_t* a_free = bars + 4;     // [2] commits  : the MMAs reading operand buffer 0 / 1 complete (one commit per MMA warp)
  uint64_t* dx_done = bars + 6;    // commit       : DX of the tile complete
  uint32_t* slot = reinterpret_cast<uint32_t*>(sm + L.slot);
  float* sdb2 = reinterpret_cast<float*>(sm + L.db2);
  const float* sb1 = reinterpret_cast<const float*>(sm + L.b1s);

  {
    const uint4* src = reinterpret_cast<const uint4*>(a.image);
    uint4* dst = reinterpret_cast<uint4*>(sm);
    for (int i = tid; i < (int)(L.image_end / 16); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (tid < 128) {     // constant parts of the x operands: K-group 2 = the ones column (hi) / zeros (lo)
    *reinterpret_cast<uint4*>(sm + L.x_hi + 2 * kBtGrp + tid * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);      // half 1.0 at d = 16
    *reinterpret_cast<uint4*>(sm + L.x_lo + 2 * kBtGrp + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid < 32) sdb2[tid] = 0.0f;
  if (tid == 0) {
    tc::mbar_init(xg_full, 256); tc::mbar_init(d_ready, 1); tc::mbar_init(d_read, 256);
    tc::mbar_init(a_ready, 256); tc::mbar_init(a_free, 3); tc::mbar_init(a_free + 1, 3); tc::mbar_init(dx_done, 1);
  }
  if (warp == 0) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  const uint32_t tD1 = tmem, tD3 = tmem + 64, tDX = tmem + 128, tW2 = tmem + 160, tW1 = tmem + 288;

  const long long rows = (long long)a.B * N;
  const int tiles = (int)((rows + 127) / 128);
  float gmax = __uint_as_float(*a.gmax_bits);
  float sG = 1.0f;
  if (gmax > 0.0f && gmax < 1e30f) { int e; frexpf(gmax, &e); sG = ldexpf(1.0f, 6 - e); }     // max |G| * sG in [32, 64)
  const float inv_sG = 1.0f / sG;

  if (warp >= 8) {
    // =============================== MMA issuers ===============================
    // tcgen05.mma issue costs ~10^2 cycles per instruction and per issuing thread and this kernel is issue-bound (69 MMAs
    // per chunk), so three warps issue side by side, each owning its accumulators:
    //   warp 8: row-wise products L1 / dh / dx      warp 9: dW2_c = h_c^T G      warp 10: dW1_c = gpre_c^T [x | 1]
    // a_free counts one commit per warp.
    if (tc::elect_one()) {
      const uint32_t id_row64 = tc::make_idesc_f16(128, 64);
      const uint32_t id_dx = tc::make_idesc_f16(128, 16) | (1u << 16);      // B = W1 (the L1 operand) read MN-major: (d, j)
      const uint32_t id_w2 = tc::make_idesc_f16(64, NP) | (1u << 15) | (1u << 16);
      const uint32_t id_w1 = tc::make_idesc_f16(64, 24) | (1u << 15) | (1u << 16);
      const uint32_t s_xh = tc::smem_u32(sm + L.x_hi), s_xl = tc::smem_u32(sm + L.x_lo);
      const uint32_t s_gh = tc::smem_u32(sm + L.g_hi), s_gl = tc::smem_u32(sm + L.g_lo);
      const uint32_t s_hh = tc::smem_u32(sm + L.hc_hi), s_hl = tc::smem_u32(sm + L.hc_lo);
      const uint32_t s_ph = tc::smem_u32(sm + L.gp_hi), s_pl = tc::smem_u32(sm + L.gp_lo);
      const uint32_t s_w1h = tc::smem_u32(sm + L.w1_hi), s_w1l = tc::smem_u32(sm + L.w1_lo);
      const uint32_t s_w2h = tc::smem_u32(sm + L.w2_hi), s_w2l = tc::smem_u32(sm + L.w2_lo);
      auto issue_l1_dh = [&](int c) {
        // L1: A = x (K = 16), B = W1 rows [c*64, c*64+64)
        bt_mma3(tD1, tc::make_desc(s_xh, kBtGrp, 128), tc::make_desc(s_xl, kBtGrp, 128),
                tc::make_desc(s_w1h + c * 64 * 16, H * 16, 128), tc::make_desc(s_w1l + c * 64 * 16, H * 16, 128), id_row64, false);
        // dh: A = G (K = NP), B = W2 rows [c*64, ...)
        for (int ks = 0; ks < NP / 16; ++ks)
          bt_mma3(tD3, tc::make_desc(s_gh + ks * 2 * kBtGrp, kBtGrp, 128), tc::make_desc(s_gl + ks * 2 * kBtGrp, kBtGrp, 128),
                  tc::make_desc(s_w2h + c * 64 * 16 + ks * 2 * H * 16, H * 16, 128),
                  tc::make_desc(s_w2l + c * 64 * 16 + ks * 2 * H * 16, H * 16, 128), id_row64, ks > 0);
        tc::mma_commit(d_ready);
      };
      uint32_t n_tile = 0, n_chunk = 0;
      for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
        if (warp == 8) {
          tc::mbar_wait(xg_full, n_tile & 1u);
          tc::fence_after_sync();
          issue_l1_dh(0);
        }
        for (int c = 0; c < nch; ++c, ++n_chunk) {
          if (warp == 8) {
            tc::mbar_wait(d_read, n_chunk & 1u);
            tc::fence_after_sync();
            if (c + 1 < nch) issue_l1_dh(c + 1);
          }
          tc::mbar_wait(a_ready, n_chunk & 1u);
          tc::fence_after_sync();
          const uint32_t bo = (n_chunk & 1u) * kBtBuf;                    // operand buffer of this chunk
          if (warp == 8) {
            // dx: A = gpre_c (K = 64, K-major), B = W1 columns [c*64, c*64+64) read MN-major as (d, j)
            for (int ks = 0; ks < 4; ++ks)
              bt_mma3(tDX, tc::make_desc(s_ph + bo + ks * 2 * kBtGrp, kBtGrp, 128), tc::make_desc(s_pl + bo + ks * 2 * kBtGrp, kBtGrp, 128),
                      tc::make_desc(s_w1h + (c * 64 + ks * 16) * 16, 128, H * 16), tc::make_desc(s_w1l + (c * 64 + ks * 16) * 16, 128, H * 16),
                      id_dx, c > 0 || ks > 0);
            if (c == nch - 1) tc::mma_commit(dx_done);
          } else {
            // weight gradients: K = the 128 rows, 8 K-steps of 16 rows (256 B each), both operands MN-major
            for (int ks = 0; ks < 8; ++ks) {
              const bool acc = n_tile > 0 || ks > 0;
              if (warp == 9)
                bt_mma3(tW2 + c * 32, tc::make_desc(s_hh + bo + ks * 256, 128, kBtGrp), tc::make_desc(s_hl + bo + ks * 256, 128, kBtGrp),
                        tc::make_desc(s_gh + ks * 256, 128, kBtGrp), tc::make_desc(s_gl + ks * 256, 128, kBtGrp), id_w2, acc);
              else
                bt_mma3(tW1 + c * 32, tc::make_desc(s_ph + bo + ks * 256, 128, kBtGrp), tc::make_desc(s_pl + bo + ks * 256, 128, kBtGrp),
                        tc::make_desc(s_xh + ks * 256, 128, kBtGrp), tc::make_desc(s_xl + ks * 256, 128, kBtGrp), id_w1, acc);
            }
          }
          tc::mma_commit(a_free + (n_chunk & 1u));
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== compute threads: two per row ===============================
    // row = tid & 127 (TMEM lane), half = tid >> 7: warp w and warp w + 4 share a lane quarter and split the columns --
    // half 0 takes hidden units [0, 32) of every chunk and stages x, half 1 takes [32, 64) and stages G
    const int row = tid & 127, half = tid >> 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    unsigned char* px_h = sm + L.x_hi + row * 16; unsigned char* px_l = sm + L.x_lo + row * 16;
    unsigned char* pg_h = sm + L.g_hi + row * 16; unsigned char* pg_l = sm + L.g_lo + row * 16;
    unsigned char* ph_h = sm + L.hc_hi + half * 4 * kBtGrp + row * 16; unsigned char* ph_l = sm + L.hc_lo + half * 4 * kBtGrp + row * 16;
    unsigned char* pp_h = sm + L.gp_hi + half * 4 * kBtGrp + row * 16; unsigned char* pp_l = sm + L.gp_lo + half * 4 * kBtGrp + row * 16;
    uint32_t n_tile = 0, n_chunk = 0;
    float db2acc[32];                                  // half 1: this thread's rows' share of db2 = sum over rows of G, reduced once per CTA
#pragma unroll
    for (int i = 0; i < 32; ++i) db2acc[i] = 0.0f;
    // every MMA issued for chunks < k complete: the last users of both operand buffers committed
    auto wait_all_mmas = [&](uint32_t k) {
      tc::mbar_wait(a_free + ((k - 1) & 1u), ((k - 1) >> 1) & 1u);
      if (k >= 2) tc::mbar_wait(a_free + ((k - 2) & 1u), ((k - 2) >> 1) & 1u);
      tc::fence_after_sync();
    };
    for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++n_tile) {
      const long long r = (long long)tile * 128 + row;
      const bool live = r < rows;
      const int b = live ? (int)(r / N) : 0, n = live ? (int)(r - (long long)b * N) : 0;
      // this thread's half of the tile's rows: x (half 0) or G (half 1), loaded into registers BEFORE waiting for the
      // previous tile's MMAs (they still read the x / G operands), so the global-load latency hides behind that drain
      float rv[32];
      {
        const float* src = half == 0 ? a.x + (long long)b * a.x_seq_stride + (long long)n * din : a.g + r * dout;
        const int w = half == 0 ? din : dout;
        if (live && (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * q < w) v4 = __ldg(reinterpret_cast<const float4*>(src) + q);
            rv[4 * q] = v4.x; rv[4 * q + 1] = v4.y; rv[4 * q + 2] = v4.z; rv[4 * q + 3] = v4.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) rv[i] = (live && i < w) ? __ldg(src + i) : 0.0f;
        }
      }
      if (n_tile > 0) { wait_all_mmas(n_chunk); }
      if (half == 0) {
#pragma unroll
        for (int kg = 0; kg < 2; ++kg) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = rv[kg * 8 + i];
          uint4 hi, lo;
          tc::split8_f16(v, hi, lo);
          *reinterpret_cast<uint4*>(px_h + kg * kBtGrp) = hi; *reinterpret_cast<uint4*>(px_l + kg * kBtGrp) = lo;
        }
      } else {
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          if (kg < NP / 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int o = kg * 8 + i;
              const float gv = rv[o];
              v[i] = gv * sG;
              db2acc[o] += gv;                                              // db2 = sum over rows of G
            }
            uint4 hi, lo;
            tc::split8_f16(v, hi, lo);
            *reinterpret_cast<uint4*>(pg_h + kg * kBtGrp) = hi; *reinterpret_cast<uint4*>(pg_l + kg * kBtGrp) = lo;
          }
        }
      }
      tc::fence_proxy_async();
      tc::fence_before_sync();
      tc::mbar_arrive(xg_full);

      for (int c = 0; c < nch; ++c, ++n_chunk) {
        tc::mbar_wait(d_ready, n_chunk & 1u);
        tc::fence_after_sync();
        uint4 hh[4], hl[4], gh[4], gl[4];
        {
          float v1[32], v3[32];
          tc::tmem_ld32(tD1 + lane_base + half * 32, v1);
          tc::tmem_ld32(tD3 + lane_base + half * 32, v3);
          tc::fence_before_sync();
          tc::mbar_arrive(d_read);                                             // D1 / D3 may be overwritten
          const float* bias = sb1 + c * kBtChunk + half * 32;
#pragma unroll
          for (int kg = 0; kg < 4; ++kg) {
            float hv[8], gv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float pre = v1[kg * 8 + i] + bias[kg * 8 + i];             // 16 * (x W1 + b1)
              hv[i] = fmaxf(pre, 0.0f);
              gv[i] = pre > 0.0f ? v3[kg * 8 + i] * (1.0f / kBtW2Scale) : 0.0f;   // sG * gpre
            }
            tc::split8_f16(hv, hh[kg], hl[kg]);
            tc::split8_f16(gv, gh[kg], gl[kg]);
          }
        }
        const uint32_t buf = n_chunk & 1u, use = n_chunk >> 1;
        if (use > 0) { tc::mbar_wait(a_free + buf, (use - 1) & 1u); tc::fence_after_sync(); }   // previous user of this buffer done
        const uint32_t bo = buf * kBtBuf;
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          *reinterpret_cast<uint4*>(ph_h + bo + kg * kBtGrp) = hh[kg]; *reinterpret_cast<uint4*>(ph_l + bo + kg * kBtGrp) = hl[kg];
          *reinterpret_cast<uint4*>(pp_h + bo + kg * kBtGrp) = gh[kg]; *reinterpret_cast<uint4*>(pp_l + bo + kg * kBtGrp) = gl[kg];
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        tc::mbar_arrive(a_ready);
      }

      // ---- gx = gpre W1^T  (half 0 writes dims [0, 8), half 1 dims [8, 16))
      tc::mbar_wait(dx_done, n_tile & 1u);
      tc::fence_after_sync();
      float dx[16];
      tc::tmem_ld16(tDX + lane_base, dx);
      tc::fence_before_sync();
      if (live) {
        const float un = inv_sG * (1.0f / kBtW1Scale);                     // DX = 16 sG (gpre W1^T)
        const int d0 = half * 8, d1 = min(din, d0 + 8);
        if (a.scatter != nullptr) {
          const int parent = a.scatter[(long long)b * a.scatter_seq_stride + n];
          float* dst = a.gx + ((long long)b * N + parent) * din;
          if (din == 16) {                                                 // two 128-bit reductions per thread instead of eight scalar ones
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (half ? dx[8 + k] : dx[k]) * un;
            atomicAdd(reinterpret_cast<float4*>(dst + d0), make_float4(v[0], v[1], v[2], v[3]));
            atomicAdd(reinterpret_cast<float4*>(dst + d0 + 4), make_float4(v[4], v[5], v[6], v[7]));
          } else {
#pragma unroll
            for (int d = 0; d < 16; ++d) if (d >= d0 && d < d1) atomicAdd(dst + d, dx[d] * un);     // static indices: dx stays in registers
          }
        } else {
          float* dst = a.gx + r * din;
          if (din == 16) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (half ? dx[8 + k] : dx[k]) * un;
            reinterpret_cast<float4*>(dst + d0)[0] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(dst + d0)[1] = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int d = 0; d < 16; ++d) if (d >= d0 && d < d1) dst[d] = dx[d] * un;
          }
        }
      }
    }

    if (half == 1) {
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        const float s = warp_sum(db2acc[o]);
        if (lane == 0 && o < dout && s != 0.0f) atomicAdd(&sdb2[o], s);
      }
    }
    // ---- flush the weight-gradient accumulators (all MMAs of this CTA complete)
    if (n_chunk > 0) wait_all_mmas(n_chunk);
    if (n_chunk > 0) {
      const bool owner = (tid & 31) < 16 && half == 0;              // M = 64 accumulator: row r lives in lane (r/16)*32 + r%16
      const int jl = ((tid >> 5) & 3) * 16 + (tid & 15);
      const float un2 = inv_sG * (1.0f / kBtW1Scale);
      for (int c = 0; c < nch; ++c) {
        const int j = c * kBtChunk + jl;
        float w2[32], w1[32];
        tc::tmem_ld32(tW2 + c * 32 + lane_base, w2);               // warp-collective: every lane takes part
        tc::tmem_ld32(tW1 + c * 32 + lane_base, w1);
        if (owner) {
          if (a.grads.W[1] != nullptr) for (int o = 0; o < dout; ++o) atomicAdd(a.grads.W[1] + (size_t)j * dout + o, w2[o] * un2);
          if (a.grads.W[0] != nullptr) for (int d = 0; d < din; ++d) atomicAdd(a.grads.W[0] + (size_t)d * H + j, w1[d] * inv_sG);
          if (a.grads.b[0] != nullptr) atomicAdd(a.grads.b[0] + j, w1[16] * inv_sG);
        }
      }
    }
    tc::fence_before_sync();
  }
  __syncthreads();
  if (tid < dout && a.grads.b[1] != nullptr) { const float v = sdb2[tid]; if (v != 0.0f) atomicAdd(a.grads.b[1] + tid, v); }
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

bool mlp_bwd_tc_supported(const psvo_mlp& net, long long rows) {
  static const bool disabled = getenv("PSVO_DISABLE_TC") != nullptr && getenv("PSVO_DISABLE_TC")[0] == '1';
  if (disabled) return false;
  if (net.n_hidden != 1 || net.d_in < 1 || net.d_in > 16 || net.d_out < 1 || net.d_out > 32) return false;
  const int H = net.width[0];
  if (H < 64 || H > 256 || (H % 64) != 0) return false;
  return rows >= 128LL * 64;
}

size_t mlp_bwd_tc_image_bytes(const psvo_mlp& net) { return bt_layout(net.width[0], net.d_out <= 16 ? 16 : 32).image_end; }

int launch_mlp_bwd_tc_stage(const psvo_mlp& net, unsigned char* image, cudaStream_t st) {
  ProfScope ps_(PK_BWD, st);
  mlp_bwd_tc_stage_kernel<<<1, 512, 0, st>>>(net, net.d_out <= 16 ? 16 : 32, image);
  PSVO_LAUNCH_CHECK();
  return 0;
}

// gmax_bits: one unsigned int of workspace (overwritten).  gx is overwritten (scatter == NULL) or accumulated through
// `scatter` (then the caller zero-fills it).  grads are accumulated.
int launch_mlp_bwd_tc(const psvo_mlp& net, const unsigned char* image, const float* x, long long x_seq_stride, const float* g,
                      int B, int N, float* gx, const int* scatter, long long scatter_seq_stride, const psvo_mlp_grad& grads,
                      unsigned int* gmax_bits, cudaStream_t st, bool gmax_ready) {
  const long long rows = (long long)B * N;
  PSVO_REQUIRE(mlp_bwd_tc_supported(net, rows), "tensor-core MLP reverse mode: unsupported shape");
  const int NP = net.d_out <= 16 ? 16 : 32;
  const BtSmem L = bt_layout(net.width[0], NP);
  const size_t smem = L.total + 128;
  PSVO_REQUIRE(smem <= 227 * 1024, "tensor-core MLP reverse mode: %zu B of shared memory", smem);
  if (!gmax_ready) {                     // the caller's kernel that wrote G did not track max |G|
    PSVO_CUDA(cudaMemsetAsync(gmax_bits, 0, sizeof(unsigned int), st));
    ProfScope ps_(PK_BWD, st);
    const long long n = rows * net.d_out;
    long long grid = (n + 255) / 256; if (grid > 4LL * num_sms()) grid = 4LL * num_sms();
    absmax_kernel<<<(unsigned)grid, 256, 0, st>>>(g, n, gmax_bits);
    PSVO_LAUNCH_CHECK();
  }
  BtArgs a;
  a.image = image; a.x = x; a.x_seq_stride = x_seq_stride; a.g = g; a.gmax_bits = gmax_bits; a.gx = gx;
  a.scatter = scatter; a.scatter_seq_stride = scatter_seq_stride; a.grads = grads;
  a.B = B; a.N = N; a.H = net.width[0]; a.NP = NP; a.d_in = net.d_in; a.d_out = net.d_out;
  const long long tiles = (rows + 127) / 128;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  PSVO_DYN_SMEM(mlp_bwd_tc_kernel, 227 * 1024);
  ProfScope ps_(PK_BWD, st);
  mlp_bwd_tc_kernel<<<grid, 352, smem, st>>>(a);
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo

using namespace psvo;

extern "C" {

PSVO_API int psvo_mlp_backward_tc_supported(const psvo_mlp* net, int64_t rows) {
  return net != nullptr && rows < (1LL << 31) && mlp_bwd_tc_supported(*net, rows) ? 1 : 0;
}

PSVO_API size_t psvo_mlp_backward_tc_workspace_bytes(const psvo_mlp* net) {
  if (net == nullptr || net->n_hidden != 1) return 0;
  return align_up(mlp_bwd_tc_image_bytes(*net), 256) + 256;
}

PSVO_API int psvo_mlp_backward_tc(const psvo_mlp* net, const float* x, const float* g_mu, int64_t rows, float* g_x,
                                  psvo_mlp_grad* grads, void* ws, size_t ws_bytes, psvo_stream_t stream) {
  PSVO_REQUIRE(net && x && g_mu && g_x && grads && ws, "psvo_mlp_backward_tc: null pointer");
  PSVO_REQUIRE(psvo_mlp_backward_tc_supported(net, rows), "psvo_mlp_backward_tc: unsupported shape (one hidden layer, width a multiple of 64 up to 256, d_in <= 16, d_out <= 32, rows >= 8192)");
  PSVO_REQUIRE(ws_bytes >= psvo_mlp_backward_tc_workspace_bytes(net), "psvo_mlp_backward_tc: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* image = static_cast<unsigned char*>(ws);
  unsigned int* gmax = reinterpret_cast<unsigned int*>(image + align_up(mlp_bwd_tc_image_bytes(*net), 256));
  PSVO_CUDA(cudaMemsetAsync(image, 0, mlp_bwd_tc_image_bytes(*net), st));
  if (int e = launch_mlp_bwd_tc_stage(*net, image, st)) return e;
  return launch_mlp_bwd_tc(*net, image, x, 0, g_mu, 1, (int)rows, g_x, nullptr, 0, *grads, gmax, st);
}

}  // extern "C"
