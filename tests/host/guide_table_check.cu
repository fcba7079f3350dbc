// guide_table_check.cu -- host-side check of the bracketed inverse-CDF search used by smc_resample_kernel
// (psvo_b200/csrc/smc_forward.cu, resample_body): for random and degenerate weight vectors the guide table
// G[k] = #{j : C[j] <= B_k} built from each particle's first slice kf(j) (float estimate + exact integer fix-up) brackets
// every child's search so that the result equals the plain upper bound over the fixed-point CDF (SVO.get_resample_idx,
// src/SMC/SVO.py:266-300 -- Categorical.sample restated as an inverse-CDF draw).  Uses the very same PSVO_HD functions as
// the kernels (fixed_weight, draw_threshold, upper_bound_u64).  Test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cstdint>
#include "psvo_math.cuh"
using namespace psvo;
int main() {
  srand(7);
  long bad = 0, probes = 0, children = 0;
  for (int trial = 0; trial < 60; ++trial) {
    const int Ns[] = {256, 1000, 4096, 300, 2048};
    const int N = Ns[trial % 5];
    int kbits = 0; while ((1 << kbits) < N) ++kbits;
    const int K = 1 << kbits, gs = 24 - kbits;
    std::vector<float> lw(N);
    const float spread = (trial % 3 == 0) ? 1.0f : (trial % 3 == 1 ? 30.0f : 200.0f);
    for (auto& v : lw) v = -spread * (rand() / (float)RAND_MAX) * (rand() / (float)RAND_MAX);
    if (trial % 7 == 0) for (int i = 0; i < N / 2; ++i) lw[rand() % N] = -1000.0f;
    float mx = -INFINITY; for (float v : lw) mx = fmaxf(mx, v);
    std::vector<uint64_t> C(N); uint64_t acc = 0;
    for (int i = 0; i < N; ++i) { acc += fixed_weight(lw[i], mx); C[i] = acc; }
    const uint64_t total = acc;
    std::vector<int> G(K + 2, 0);
    const float scale = (float)K / (float)total;
    for (int j = 0; j < N; ++j) {
      const uint64_t c = C[j];
      int kf = (int)((float)c * scale);
      kf = kf < 0 ? 0 : (kf > K ? K : kf);
      int fix = 0;
      while (kf > 0 && draw_threshold(total, (uint32_t)(kf - 1) << gs) >= c) { --kf; ++fix; }
      while (kf < K && draw_threshold(total, (uint32_t)kf << gs) < c) { ++kf; ++fix; }
      if (fix > 3) printf("fix %d\n", fix);
      G[kf]++;
    }
    for (int k = 1; k <= K; ++k) G[k] += G[k - 1];
    if (G[K] != N) { printf("G[K]=%d N=%d\n", G[K], N); ++bad; }
    for (int n = 0; n < 20000; ++n) {
      uint32_t u24 = (uint32_t)(((uint64_t)rand() << 9) ^ rand()) & 0xFFFFFFu;
      if (n < 4) u24 = n == 0 ? 0 : (n == 1 ? 0xFFFFFFu : (n == 2 ? 1u << gs : (1u << gs) - 1));
      const uint64_t thr = draw_threshold(total, u24);
      int ref = upper_bound_u64(C.data(), N, thr);
      int kb = (int)(u24 >> gs); if (kb > K - 1) kb = K - 1;
      int lo = G[kb], hi = G[kb + 1];
      while (lo < hi) { const int mid = (lo + hi) >> 1; ++probes; if (C[mid] <= thr) lo = mid + 1; else hi = mid; }
      const int got = lo < N ? lo : N - 1;
      ++children;
      if (got != ref) { if (bad < 10) printf("mismatch N=%d u24=%u got %d ref %d\n", N, u24, got, ref); ++bad; }
    }
  }
  printf("bad=%ld children=%ld probes/child=%.2f\n", bad, children, (double)probes / children);
  return bad != 0;
}
