// mlp_tile.cuh -- MLP_transformation evaluated per particle by one thread, weights staged in shared memory.
//
// Reference: MLP_transformation.transform (src/transformation/MLP.py:48-54): ReLU Dense stack + linear mu head,
// Keras kernels [in,out].  This is the register-tiled FP32 form used by the fused SMC kernels: every lane of a
// warp reads the SAME weight (a shared-memory broadcast, LDS.128) and applies it to its own PPT particles, so
// the per-particle MLP needs no cross-thread traffic and no synchronisation.
//   L = 0 : out = b + x W                      (linear nets of config 3)
//   L = 1 : hidden unit j is formed, rectified and consumed in registers (never stored)      <- hot case
//   L >= 2: hidden activations ping-pong through per-thread shared-memory columns (general, slower)
#pragma once
#include "psvo_common.cuh"

namespace psvo {

template <int DIN, int DOUT>
struct MlpS {
  int L;
  int H[PSVO_MAX_HIDDEN];
  const float* w0t;                       // [H0][DIN]   first layer, transposed + zero padded   (L >= 1)
  const float* b0;                        // [H0]
  const float* wmid[PSVO_MAX_HIDDEN];     // [H_{k-1}][H_k] as stored (k = 1..L-1)
  const float* bmid[PSVO_MAX_HIDDEN];
  const float* wh;                        // [H_{L-1} or DIN][DOUT] head, zero padded
  const float* bh;                        // [DOUT]
};

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

// floats of shared memory mlp_stage() consumes
template <int DIN, int DOUT>
__host__ __device__ inline int mlp_smem_floats(const psvo_mlp& net) {
  const int L = net.n_hidden;
  if (L == 0) return DIN * DOUT + DOUT;
  int n = net.width[0] * DIN + round4(net.width[0]);
  for (int k = 1; k < L; ++k) n += round4(net.width[k - 1] * net.width[k]) + round4(net.width[k]);
  n += net.width[L - 1] * DOUT + DOUT;
  return n;
}

#if defined(__CUDACC__)
// Cooperative (whole block) copy of the net into shared memory; returns the next free float.
// Caller must __syncthreads() before using the staged net.
template <int DIN, int DOUT>
__device__ inline float* mlp_stage(const psvo_mlp& net, float* smem, MlpS<DIN, DOUT>& s) {
  const int L = net.n_hidden, din = net.d_in, dout = net.d_out;
  const int tid = threadIdx.x, nt = blockDim.x;
  s.L = L;
  for (int k = 0; k < PSVO_MAX_HIDDEN; ++k) s.H[k] = k < L ? net.width[k] : 0;
  float* p = smem;
  if (L >= 1) {
    const int H0 = net.width[0];
    float* w0t = p; p += H0 * DIN;
    float* b0 = p;  p += round4(H0);
    for (int i = tid; i < H0 * DIN; i += nt) {
      const int j = i / DIN, d = i - j * DIN;
      w0t[i] = d < din ? __ldg(net.W[0] + (size_t)d * H0 + j) : 0.0f;
    }
    for (int i = tid; i < H0; i += nt) b0[i] = __ldg(net.b[0] + i);
    s.w0t = w0t; s.b0 = b0;
    for (int k = 1; k < L; ++k) {
      const int hi = net.width[k - 1], ho = net.width[k];
      float* w = p; p += round4(hi * ho);
      float* b = p; p += round4(ho);
      for (int i = tid; i < hi * ho; i += nt) w[i] = __ldg(net.W[k] + i);
      for (int i = tid; i < ho; i += nt) b[i] = __ldg(net.b[k] + i);
      s.wmid[k] = w; s.bmid[k] = b;
    }
  }
  const int hin = L == 0 ? DIN : net.width[L - 1];      // rows of the staged head
  const int hin_real = L == 0 ? din : net.width[L - 1];
  float* wh = p; p += hin * DOUT;
  float* bh = p; p += DOUT;
  for (int i = tid; i < hin * DOUT; i += nt) {
    const int r = i / DOUT, o = i - r * DOUT;
    wh[i] = (r < hin_real && o < dout) ? __ldg(net.W[L] + (size_t)r * dout + o) : 0.0f;
  }
  for (int i = tid; i < DOUT; i += nt) bh[i] = i < dout ? __ldg(net.b[L] + i) : 0.0f;
  s.wh = wh; s.bh = bh;
  return p;
}

template <int N>
__device__ __forceinline__ void lds_vec(const float* __restrict__ src, float (&dst)[N]) {
  static_assert(N % 4 == 0, "vector width");
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    dst[4 * i + 0] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
  }
}

// out[p][:] = MLP(x[p][:]) for the PPT particles this thread owns.
// scratch: two ping-pong regions of Hmax * NT floats (NT = blockDim.x * PPT) when s.L >= 2, else unused.
template <int DIN, int DOUT, int PPT>
__device__ __forceinline__ void mlp_eval(const MlpS<DIN, DOUT>& s, const float (&x)[PPT][DIN],
                                         float (&out)[PPT][DOUT], float* scratch, int scratch_half) {
#pragma unroll
  for (int p = 0; p < PPT; ++p)
#pragma unroll
    for (int o = 0; o < DOUT; ++o) out[p][o] = s.bh[o];

  if (s.L == 0) {
#pragma unroll
    for (int d = 0; d < DIN; ++d) {
      float w[DOUT];
      lds_vec<DOUT>(s.wh + d * DOUT, w);
#pragma unroll
      for (int p = 0; p < PPT; ++p)
#pragma unroll
        for (int o = 0; o < DOUT; ++o) out[p][o] = fmaf(x[p][d], w[o], out[p][o]);
    }
    return;
  }

  if (s.L == 1) {
    const int H = s.H[0];
#pragma unroll 2
    for (int j = 0; j < H; ++j) {
      float w[DIN];
      lds_vec<DIN>(s.w0t + j * DIN, w);
      const float bj = s.b0[j];
      float h[PPT];
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        float a = bj;
#pragma unroll
        for (int d = 0; d < DIN; ++d) a = fmaf(x[p][d], w[d], a);
        h[p] = fmaxf(a, 0.0f);
      }
      float wo[DOUT];
      lds_vec<DOUT>(s.wh + j * DOUT, wo);
#pragma unroll
      for (int p = 0; p < PPT; ++p)
#pragma unroll
        for (int o = 0; o < DOUT; ++o) out[p][o] = fmaf(h[p], wo[o], out[p][o]);
    }
    return;
  }

  // general depth: activations of thread-particle q live at A[i * NT + q]
  const int NT = blockDim.x * PPT;
  float* A = scratch;
  float* Bf = scratch + scratch_half;
  {
    const int H0 = s.H[0];
    for (int j = 0; j < H0; ++j) {
      float w[DIN];
      lds_vec<DIN>(s.w0t + j * DIN, w);
      const float bj = s.b0[j];
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        float a = bj;
#pragma unroll
        for (int d = 0; d < DIN; ++d) a = fmaf(x[p][d], w[d], a);
        A[j * NT + p * blockDim.x + threadIdx.x] = fmaxf(a, 0.0f);
      }
    }
  }
  for (int k = 1; k < s.L; ++k) {
    const int hi = s.H[k - 1], ho = s.H[k];
    for (int j = 0; j < ho; ++j) {
      const float bj = s.bmid[k][j];
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        float a = bj;
        for (int i = 0; i < hi; ++i) a = fmaf(A[i * NT + p * blockDim.x + threadIdx.x], s.wmid[k][i * ho + j], a);
        Bf[j * NT + p * blockDim.x + threadIdx.x] = fmaxf(a, 0.0f);
      }
    }
    float* t = A; A = Bf; Bf = t;
  }
  {
    const int hl = s.H[s.L - 1];
    for (int i = 0; i < hl; ++i) {
      float wo[DOUT];
      lds_vec<DOUT>(s.wh + i * DOUT, wo);
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        const float h = A[i * NT + p * blockDim.x + threadIdx.x];
#pragma unroll
        for (int o = 0; o < DOUT; ++o) out[p][o] = fmaf(h, wo[o], out[p][o]);
      }
    }
  }
}
#endif  // __CUDACC__

inline int mlp_max_width(const psvo_mlp& net) {
  int m = 0;
  for (int k = 0; k < net.n_hidden; ++k) m = net.width[k] > m ? net.width[k] : m;
  return m;
}

}  // namespace psvo
