// mlp_bwd_tile.cuh -- reverse mode of the per-particle MLP (mlp_tile.cuh) for one tile of P = blockDim.x particles.
//
// Reference semantics: TF autodiff through MLP_transformation.transform (src/transformation/MLP.py:48-54).
// Each thread owns one particle and back-propagates its upstream gradient to its input in registers
// (phase A).  The weight gradients are sums over particles, i.e. small GEMMs  dW2 = H^T G,  dW1 = X^T Gpre:
// phase A parks h_j and gpre_j of a chunk of 32 hidden units in shared memory, then in phase B the block
// re-partitions itself as (unit, particle segment) and every thread accumulates the (DIN + DOUT + 1) gradient
// entries of its unit over its segment in registers, adding them to the CTA's shared-memory gradient
// accumulator (same layout as the staged weights) with shared-memory atomics.  The accumulator is flushed to
// global memory once per CTA per launch.
//   L = 0 (linear) and L = 1 (one hidden layer) are supported.
#pragma once
#include "mlp_tile.cuh"

namespace psvo {

constexpr int kBwdChunk = 32;     // hidden units per chunk == lanes of a warp

// floats of staging shared memory mlp_bwd_tile needs for a block of P threads
template <int DIN, int DOUT>
__host__ __device__ inline int mlp_bwd_stage_floats(int P) {
  return 2 * kBwdChunk * (P + 1) + P * DIN + P * DOUT;
}

#if defined(__CUDACC__)
// dW: shared-memory accumulator laid out exactly like the staged net (mlp_stage): index with the same offsets.
template <int DIN, int DOUT>
struct MlpGradS {
  float* w0t; float* b0; float* wh; float* bh;
};

template <int DIN, int DOUT>
__device__ inline float* mlp_grad_carve(const psvo_mlp& net, float* smem, MlpGradS<DIN, DOUT>& g) {
  const int L = net.n_hidden;
  float* p = smem;
  if (L >= 1) {
    g.w0t = p; p += net.width[0] * DIN;
    g.b0 = p;  p += round4(net.width[0]);
  } else { g.w0t = nullptr; g.b0 = nullptr; }
  const int hin = L == 0 ? DIN : net.width[L - 1];
  g.wh = p; p += hin * DOUT;
  g.bh = p; p += DOUT;
  return p;
}

// flush the CTA's accumulator into the caller's gradient buffers (Keras layout), one atomic per parameter
template <int DIN, int DOUT>
__device__ inline void mlp_grad_flush(const psvo_mlp& net, const MlpGradS<DIN, DOUT>& g, const psvo_mlp_grad& out) {
  const int L = net.n_hidden, din = net.d_in, dout = net.d_out;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (L >= 1) {
    const int H = net.width[0];
    if (out.W[0] != nullptr)
      for (int i = tid; i < H * din; i += nt) {          // out.W[0][d][j] <- w0t[j][d]
        const int d = i / H, j = i - d * H;
        const float v = g.w0t[j * DIN + d];
        if (v != 0.0f) atomicAdd(out.W[0] + i, v);
      }
    if (out.b[0] != nullptr)
      for (int i = tid; i < H; i += nt) { const float v = g.b0[i]; if (v != 0.0f) atomicAdd(out.b[0] + i, v); }
  }
  const int hin = L == 0 ? din : net.width[L - 1];
  if (out.W[L] != nullptr)
    for (int i = tid; i < hin * dout; i += nt) {
      const int r = i / dout, o = i - r * dout;
      const float v = g.wh[r * DOUT + o];
      if (v != 0.0f) atomicAdd(out.W[L] + i, v);
    }
  if (out.b[L] != nullptr)
    for (int i = tid; i < dout; i += nt) { const float v = g.bh[i]; if (v != 0.0f) atomicAdd(out.b[L] + i, v); }
}

// Collective over the block.  x / gout: this thread's particle (zero gout for dead lanes).  gx is overwritten.
template <int DIN, int DOUT>
__device__ __forceinline__ void mlp_bwd_tile(const MlpS<DIN, DOUT>& s, const MlpGradS<DIN, DOUT>& g,
                                             const float (&x)[DIN], const float (&gout)[DOUT], float (&gx)[DIN],
                                             float* stage) {
  const int P = blockDim.x, tid = threadIdx.x;
  float* sH = stage;                                   // [kBwdChunk][P+1]
  float* sGp = sH + kBwdChunk * (P + 1);               // [kBwdChunk][P+1]
  float* sX = sGp + kBwdChunk * (P + 1);               // [P][DIN]
  float* sG = sX + P * DIN;                            // [P][DOUT]
  const int lane = tid & 31, warp = tid >> 5, nwarp = P >> 5;

  __syncthreads();                                     // staging area free (previous use finished)
#pragma unroll
  for (int d = 0; d < DIN; ++d) { sX[tid * DIN + d] = x[d]; gx[d] = 0.0f; }
#pragma unroll
  for (int o = 0; o < DOUT; ++o) sG[tid * DOUT + o] = gout[o];
  __syncthreads();

  // head bias gradient: sum_p gout[p][o]
  for (int o = tid; o < DOUT; o += P) {
    float acc = 0.0f;
    for (int p = 0; p < P; ++p) acc += sG[p * DOUT + o];
    g.bh[o] += acc;                                    // one owner thread per o
  }

  if (s.L == 0) {
    // gx = W gout ; dW[d][o] += sum_p x[p][d] gout[p][o]
#pragma unroll
    for (int d = 0; d < DIN; ++d) {
      float w[DOUT];
      lds_vec<DOUT>(s.wh + d * DOUT, w);
      float acc = 0.0f;
#pragma unroll
      for (int o = 0; o < DOUT; ++o) acc = fmaf(w[o], gout[o], acc);
      gx[d] = acc;
    }
    // (row d, segment): DIN rows x nseg segments
    const int nseg = P / DIN > 0 ? P / DIN : 1;
    const int d = tid % DIN, seg = tid / DIN;
    if (seg < nseg) {
      const int per = (P + nseg - 1) / nseg;
      const int p0 = seg * per, p1 = min(P, p0 + per);
      float acc[DOUT];
#pragma unroll
      for (int o = 0; o < DOUT; ++o) acc[o] = 0.0f;
      for (int p = p0; p < p1; ++p) {
        const float xv = sX[p * DIN + d];
        float gv[DOUT];
        lds_vec<DOUT>(sG + p * DOUT, gv);
#pragma unroll
        for (int o = 0; o < DOUT; ++o) acc[o] = fmaf(xv, gv[o], acc[o]);
      }
#pragma unroll
      for (int o = 0; o < DOUT; ++o) atomicAdd(g.wh + d * DOUT + o, acc[o]);
    }
    return;
  }

  // one hidden layer
  const int H = s.H[0];
  for (int c0 = 0; c0 < H; c0 += kBwdChunk) {
    const int nj = min(kBwdChunk, H - c0);
    // ---- phase A: per particle, units c0 .. c0+nj
    for (int jj = 0; jj < nj; ++jj) {
      const int j = c0 + jj;
      float w[DIN];
      lds_vec<DIN>(s.w0t + j * DIN, w);
      float pre = s.b0[j];
#pragma unroll
      for (int d = 0; d < DIN; ++d) pre = fmaf(x[d], w[d], pre);
      float wo[DOUT];
      lds_vec<DOUT>(s.wh + j * DOUT, wo);
      float gh = 0.0f;
#pragma unroll
      for (int o = 0; o < DOUT; ++o) gh = fmaf(wo[o], gout[o], gh);
      const float gp = pre > 0.0f ? gh : 0.0f;
#pragma unroll
      for (int d = 0; d < DIN; ++d) gx[d] = fmaf(w[d], gp, gx[d]);
      sH[jj * (P + 1) + tid] = fmaxf(pre, 0.0f);
      sGp[jj * (P + 1) + tid] = gp;
    }
    __syncthreads();
    // ---- phase B: lane = unit within the chunk, warp = particle segment
    if (lane < nj) {
      const int j = c0 + lane;
      const int per = (P + nwarp - 1) / nwarp;
      const int p0 = warp * per, p1 = min(P, p0 + per);
      float a2[DOUT], a1[DIN], ab = 0.0f;
#pragma unroll
      for (int o = 0; o < DOUT; ++o) a2[o] = 0.0f;
#pragma unroll
      for (int d = 0; d < DIN; ++d) a1[d] = 0.0f;
      for (int p = p0; p < p1; ++p) {
        const float h = sH[lane * (P + 1) + p];
        const float gp = sGp[lane * (P + 1) + p];
        float gv[DOUT], xv[DIN];
        lds_vec<DOUT>(sG + p * DOUT, gv);
        lds_vec<DIN>(sX + p * DIN, xv);
#pragma unroll
        for (int o = 0; o < DOUT; ++o) a2[o] = fmaf(h, gv[o], a2[o]);
#pragma unroll
        for (int d = 0; d < DIN; ++d) a1[d] = fmaf(gp, xv[d], a1[d]);
        ab += gp;
      }
#pragma unroll
      for (int o = 0; o < DOUT; ++o) atomicAdd(g.wh + j * DOUT + o, a2[o]);
#pragma unroll
      for (int d = 0; d < DIN; ++d) atomicAdd(g.w0t + j * DIN + d, a1[d]);
      atomicAdd(g.b0 + j, ab);
    }
    __syncthreads();
  }
}
#endif  // __CUDACC__

}  // namespace psvo
