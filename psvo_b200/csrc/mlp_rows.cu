// mlp_rows.cu -- MLP_transformation.transform on a (rows, d_in) matrix, forward and reverse mode.
//
// Reference: src/transformation/MLP.py:48-54.  Used for the networks that are evaluated once per sequence and
// time step rather than per particle (q0, q2, BSim_q_init, BSim_q2: B or B*T rows), and as the stand-alone
// psvo_mlp_forward / psvo_mlp_backward operators behind MLP_transformation.transform.
// A CTA walks row tiles; activations of the tile sit in shared memory; weights stream from L2 with
// consecutive output units on consecutive lanes (coalesced).  Reverse mode keeps the CTA's weight-gradient
// partial sums in shared memory across all its tiles and flushes them with one atomicAdd per parameter.
#include "smc_kernels.h"

namespace psvo {

constexpr int kRowTile = 16;

struct MlpShape { int L, dims[PSVO_MAX_HIDDEN + 2], maxw, sum_in, n_param; };

static MlpShape shape_of(const psvo_mlp& net) {
  MlpShape s;
  s.L = net.n_hidden;
  s.dims[0] = net.d_in;
  for (int k = 0; k < s.L; ++k) s.dims[k + 1] = net.width[k];
  s.dims[s.L + 1] = net.d_out;
  s.maxw = 0; s.sum_in = 0; s.n_param = 0;
  for (int k = 0; k <= s.L + 1; ++k) s.maxw = s.dims[k] > s.maxw ? s.dims[k] : s.maxw;
  for (int k = 0; k <= s.L; ++k) { s.sum_in += s.dims[k]; s.n_param += s.dims[k] * s.dims[k + 1] + s.dims[k + 1]; }
  return s;
}

__global__ void __launch_bounds__(256) mlp_rows_forward_kernel(psvo_mlp net, MlpShape s, const float* __restrict__ x,
                                                               long long rows, float* __restrict__ mu) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm;
  float* Bf = sm + kRowTile * s.maxw;
  const int tid = threadIdx.x, nt = blockDim.x;
  const long long tiles = (rows + kRowTile - 1) / kRowTile;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r0 = tile * kRowTile;
    const int din = s.dims[0];
    __syncthreads();
    for (int i = tid; i < kRowTile * din; i += nt) {
      const int r = i / din, c = i - r * din;
      A[r * s.maxw + c] = (r0 + r < rows) ? x[(r0 + r) * din + c] : 0.0f;
    }
    __syncthreads();
    float* in = A; float* out = Bf;
    for (int k = 0; k <= s.L; ++k) {
      const int ni = s.dims[k], no = s.dims[k + 1];
      const float* __restrict__ W = net.W[k];
      const float* __restrict__ bb = net.b[k];
      for (int i = tid; i < kRowTile * no; i += nt) {
        const int r = i / no, j = i - r * no;
        float acc = __ldg(bb + j);
        const float* arow = in + r * s.maxw;
        for (int c = 0; c < ni; ++c) acc = fmaf(arow[c], __ldg(W + (size_t)c * no + j), acc);
        out[r * s.maxw + j] = k < s.L ? fmaxf(acc, 0.0f) : acc;
      }
      __syncthreads();
      float* t = in; in = out; out = t;
    }
    const int dout = s.dims[s.L + 1];
    for (int i = tid; i < kRowTile * dout; i += nt) {
      const int r = i / dout, c = i - r * dout;
      if (r0 + r < rows) mu[(r0 + r) * dout + c] = in[r * s.maxw + c];
    }
  }
}

int launch_mlp_rows_forward(const psvo_mlp& net, const float* x, int64_t rows, float* mu, cudaStream_t st) {
  if (rows <= 0) return 0;
  PSVO_REQUIRE(net.n_hidden >= 0 && net.n_hidden <= PSVO_MAX_HIDDEN, "n_hidden=%d out of range", net.n_hidden);
  const MlpShape s = shape_of(net);
  const size_t smem = 2 * (size_t)kRowTile * s.maxw * sizeof(float);
  PSVO_REQUIRE(smem <= 200 * 1024, "MLP layer width %d too large", s.maxw);
  PSVO_DYN_SMEM(mlp_rows_forward_kernel, smem);
  const int64_t tiles = (rows + kRowTile - 1) / kRowTile;
  const int grid = (int)(tiles < 4 * (int64_t)num_sms() ? tiles : 4 * (int64_t)num_sms());
  ProfScope ps_(PK_ROWS, st);
  mlp_rows_forward_kernel<<<grid, 256, smem, st>>>(net, s, x, (long long)rows, mu);
  PSVO_LAUNCH_CHECK();
  return 0;
}

// reverse mode.  shared memory: Act (all layer inputs of the tile) | G ping | G pong | gP (parameter partials)
__global__ void __launch_bounds__(256) mlp_rows_backward_kernel(psvo_mlp net, MlpShape s, psvo_mlp_grad gr,
                                                                const float* __restrict__ x,
                                                                const float* __restrict__ g_mu, long long rows,
                                                                float* __restrict__ g_x) {
  extern __shared__ __align__(16) float sm[];
  float* Act = sm;                                   // [kRowTile][sum_in]
  float* G0 = Act + kRowTile * s.sum_in;             // [kRowTile][maxw]
  float* G1 = G0 + kRowTile * s.maxw;
  float* gP = G1 + kRowTile * s.maxw;                // [n_param] : per layer W then b
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < s.n_param; i += nt) gP[i] = 0.0f;
  const long long tiles = (rows + kRowTile - 1) / kRowTile;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long r0 = tile * kRowTile;
    __syncthreads();
    {
      const int din = s.dims[0];
      for (int i = tid; i < kRowTile * din; i += nt) {
        const int r = i / din, c = i - r * din;
        Act[r * s.sum_in + c] = (r0 + r < rows) ? x[(r0 + r) * din + c] : 0.0f;
      }
    }
    __syncthreads();
    int off = 0;
    for (int k = 0; k < s.L; ++k) {                  // recompute hidden activations
      const int ni = s.dims[k], no = s.dims[k + 1];
      const float* __restrict__ W = net.W[k];
      const float* __restrict__ bb = net.b[k];
      for (int i = tid; i < kRowTile * no; i += nt) {
        const int r = i / no, j = i - r * no;
        float acc = __ldg(bb + j);
        const float* arow = Act + r * s.sum_in + off;
        for (int c = 0; c < ni; ++c) acc = fmaf(arow[c], __ldg(W + (size_t)c * no + j), acc);
        Act[r * s.sum_in + off + ni + j] = fmaxf(acc, 0.0f);
      }
      off += ni;
      __syncthreads();
    }
    {
      const int dout = s.dims[s.L + 1];
      for (int i = tid; i < kRowTile * dout; i += nt) {
        const int r = i / dout, c = i - r * dout;
        G0[r * s.maxw + c] = (r0 + r < rows) ? g_mu[(r0 + r) * dout + c] : 0.0f;
      }
    }
    __syncthreads();
    float* G = G0; float* Gn = G1;
    int poff = s.n_param;
    for (int k = s.L; k >= 0; --k) {
      const int ni = s.dims[k], no = s.dims[k + 1];
      poff -= ni * no + no;
      const float* __restrict__ W = net.W[k];
      const float* act = Act + off;                  // layer input, row stride sum_in   (off = offset of layer k input)
      for (int i = tid; i < ni * no; i += nt) {      // dW[c][j] += sum_r act[r][c] G[r][j]
        const int c = i / no, j = i - c * no;
        float acc = 0.0f;
#pragma unroll
        for (int r = 0; r < kRowTile; ++r) acc = fmaf(act[r * s.sum_in + c], G[r * s.maxw + j], acc);
        gP[poff + i] += acc;
      }
      for (int j = tid; j < no; j += nt) {
        float acc = 0.0f;
#pragma unroll
        for (int r = 0; r < kRowTile; ++r) acc += G[r * s.maxw + j];
        gP[poff + ni * no + j] += acc;
      }
      if (k > 0 || g_x != nullptr) {
        for (int i = tid; i < kRowTile * ni; i += nt) {   // g_in[r][c] = sum_j W[c][j] G[r][j]  (* relu')
          const int r = i / ni, c = i - r * ni;
          float acc = 0.0f;
          const float* grow = G + r * s.maxw;
          for (int j = 0; j < no; ++j) acc = fmaf(__ldg(W + (size_t)c * no + j), grow[j], acc);
          if (k > 0 && !(act[r * s.sum_in + c] > 0.0f)) acc = 0.0f;
          Gn[r * s.maxw + c] = acc;
        }
      }
      __syncthreads();
      float* t = G; G = Gn; Gn = t;
      if (k > 0) off -= s.dims[k - 1];
    }
    if (g_x != nullptr) {
      const int din = s.dims[0];
      for (int i = tid; i < kRowTile * din; i += nt) {
        const int r = i / din, c = i - r * din;
        if (r0 + r < rows) g_x[(r0 + r) * din + c] = G[r * s.maxw + c];
      }
    }
  }
  __syncthreads();
  int poff = 0;
  for (int k = 0; k <= s.L; ++k) {
    const int nw = s.dims[k] * s.dims[k + 1], nb = s.dims[k + 1];
    if (gr.W[k] != nullptr) for (int i = tid; i < nw; i += nt) atomicAdd(gr.W[k] + i, gP[poff + i]);
    if (gr.b[k] != nullptr) for (int i = tid; i < nb; i += nt) atomicAdd(gr.b[k] + i, gP[poff + nw + i]);
    poff += nw + nb;
  }
}

int launch_mlp_rows_backward(const psvo_mlp& net, const float* x, const float* g_mu, int64_t rows,
                             float* g_x, const psvo_mlp_grad* grads, cudaStream_t st) {
  if (rows <= 0) return 0;
  PSVO_REQUIRE(net.n_hidden >= 0 && net.n_hidden <= PSVO_MAX_HIDDEN, "n_hidden=%d out of range", net.n_hidden);
  const MlpShape s = shape_of(net);
  const size_t smem = ((size_t)kRowTile * (s.sum_in + 2 * s.maxw) + s.n_param) * sizeof(float);
  PSVO_REQUIRE(smem <= 220 * 1024, "MLP too large for the row-backward kernel (%zu B shared memory)", smem);
  PSVO_DYN_SMEM(mlp_rows_backward_kernel, smem);
  psvo_mlp_grad gr;
  if (grads != nullptr) gr = *grads;
  else for (int k = 0; k <= PSVO_MAX_HIDDEN; ++k) { gr.W[k] = nullptr; gr.b[k] = nullptr; }
  const int64_t tiles = (rows + kRowTile - 1) / kRowTile;
  const int grid = (int)(tiles < 2 * (int64_t)num_sms() ? tiles : 2 * (int64_t)num_sms());
  ProfScope ps_(PK_ROWS, st);
  mlp_rows_backward_kernel<<<grid, 256, smem, st>>>(net, s, gr, x, g_mu, (long long)rows, g_x);
  PSVO_LAUNCH_CHECK();
  return 0;
}

}  // namespace psvo
