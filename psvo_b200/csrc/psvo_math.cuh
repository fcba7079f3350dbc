// psvo_math.cuh -- scalar arithmetic shared by every kernel (and by the host-side self-check build).
// Everything here is __host__ __device__ so tests/hostcheck can compile it with g++ and compare bits
// with the numpy oracle (oracle/psvo_oracle.py: det_exp_f32, fixed_point_weights, resample_idx).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PSVO_HD __host__ __device__ __forceinline__
#else
#define PSVO_HD inline
#endif

namespace psvo {

constexpr float kLog2Pi = 1.8378770664093453f;
constexpr float kLn2 = 0.6931471805599453f;

// ---- IEEE fp32 multiply / add that the compiler may not contract into an FMA -------------------------
PSVO_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b;
  return r;
#endif
}
PSVO_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}
PSVO_HD float bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}

// exp(x) for x <= 0 from fp32 multiply/add only; returns 0 below -23 (under the 2^-32 resolution of the
// fixed-point weights).  Bit-identical to oracle det_exp_f32.
PSVO_HD float det_exp(float x) {
  if (!(x >= -23.0f)) return 0.0f;          // also catches NaN -> weight 0
  const float t = mul_rn(x, 1.4426950408889634f);
  const float n = rintf(t);
  const float f = add_rn(t, -n);
  float p = 0.00015403530393381608f;
  p = add_rn(mul_rn(p, f), 0.0013333558146428443f);
  p = add_rn(mul_rn(p, f), 0.009618129107628477f);
  p = add_rn(mul_rn(p, f), 0.05550410866482158f);
  p = add_rn(mul_rn(p, f), 0.2402265069591007f);
  p = add_rn(mul_rn(p, f), 0.6931471805599453f);
  p = add_rn(mul_rn(p, f), 1.0f);
  const float scale = bits_to_float(((uint32_t)((int)n + 127)) << 23);
  return mul_rn(p, scale);
}

// floor(exp(l - max) * 2^32) as an integer weight (the max element maps to exactly 2^32)
PSVO_HD uint64_t fixed_weight(float l, float mx) {
  const float w = det_exp(add_rn(l, -mx));
  return (uint64_t)mul_rn(w, 4294967296.0f);
}

// threshold floor(total * floor(u*2^24) / 2^24) without 128-bit arithmetic
PSVO_HD uint64_t draw_threshold(uint64_t total, uint32_t u24) {
  const uint64_t th = total >> 24, tl = total & 0xFFFFFFull;
  return th * (uint64_t)u24 + ((tl * (uint64_t)u24) >> 24);
}
PSVO_HD uint32_t u24_from_float(float u) { return (uint32_t)mul_rn(u, 16777216.0f); }

// #{k : C[k] <= thr} over an inclusive prefix sum (upper bound); result in [0, K-1] because thr < C[K-1]
template <typename Ptr>
PSVO_HD int upper_bound_u64(Ptr C, int K, uint64_t thr) {
  int lo = 0, hi = K;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (C[mid] <= thr) lo = mid + 1; else hi = mid;
  }
  return lo < K ? lo : K - 1;
}

// ---- Philox4x32-10 counter-based generator ------------------------------------------------------------
struct Philox4 { uint32_t v[4]; };

PSVO_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

PSVO_HD Philox4 philox4x32_10(uint64_t seed, uint64_t index, uint32_t chunk, uint32_t stream_id) {
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = chunk, c3 = stream_id;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// The uniform of element `index` of stream `stream_id`: word (index & 3) of the Philox block of counter index >> 2, so four
// consecutive elements cost ONE block (the resampling kernel draws the uniforms of four consecutive children at once).
PSVO_HD uint32_t uniform_bits(uint64_t seed, uint64_t index, uint32_t stream_id) {
  return philox4x32_10(seed, index >> 2, 0u, stream_id).v[index & 3u];
}

// 24-bit uniform in [0,1)
PSVO_HD float uniform_from_bits(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

// four N(0,1) draws from one Philox block (two Box-Muller pairs).  Device: MUFU-based log / sqrt / sin / cos (the draws
// only have to be N(0,1) and a pure function of the Philox block -- psvo_fill_normal materialises exactly these values); the
// angle is taken in [-pi, pi) where the fast sine/cosine are accurate, and the radicand is clamped because the fast log2 has
// an absolute, not relative, error next to u1 = 1.  16 normals cost ~100 instructions instead of ~650.
PSVO_HD void normal4(const Philox4& r, float z[4]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float u1 = ((float)(r.v[2 * k] >> 8) + 1.0f) * 5.9604644775390625e-08f;   // (0,1]
    const float u2 = uniform_from_bits(r.v[2 * k + 1]);
    float s, c;
#if defined(__CUDA_ARCH__)
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(fmaxf(-2.0f * __logf(u1), 0.0f)));
    __sincosf(6.283185307179586f * (u2 - 0.5f), &s, &c);
#else
    const float rad = sqrtf(fmaxf(-2.0f * logf(u1), 0.0f));
    s = sinf(6.283185307179586f * (u2 - 0.5f)); c = cosf(6.283185307179586f * (u2 - 0.5f));
#endif
    z[2 * k] = rad * c;
    z[2 * k + 1] = rad * s;
  }
}

// ---- small numerics -----------------------------------------------------------------------------------
PSVO_HD float softplus(float x) { return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x))); }
PSVO_HD float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// tf_mvn.get_sigma: softplus, NaN -> 0, clamp from below
PSVO_HD float sigma_from_con(float sc, float smin) {
  float s = softplus(sc);
  if (s != s) s = 0.0f;
  return fmaxf(s, smin);
}

}  // namespace psvo
