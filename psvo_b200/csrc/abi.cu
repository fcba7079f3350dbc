// abi.cu -- extern "C" entry points of include/psvo.h for the forward filter, plus error plumbing.
#include <stdarg.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <utility>
#include <vector>
#include "smc_kernels.h"

namespace psvo {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return -1;
}

int num_sms() {
  static int cache[64] = {0};                   // per device: one process may drive several GPUs
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev] = n;
  }
  return cache[dev];
}

int ensure_dyn_smem(const void* kernel, size_t bytes) {
  if (bytes <= 40 * 1024) return 0;             // within the default limit
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> have;
  int dev = 0;
  PSVO_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t& h = have[std::make_pair(dev, kernel)];
  if (bytes > h) {
    PSVO_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    h = bytes;
  }
  return 0;
}

// ---- measurement hooks
namespace {
struct Prof {
  std::mutex mu;
  bool on = false;
  long long launches[PSVO_PROF_KINDS] = {0};
  struct Pair { cudaEvent_t a, b; int kind; };
  std::vector<Pair> open_pairs;      // begun, end not yet recorded (LIFO per thread is enough here)
  std::vector<Pair> done;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
} g_prof;
}  // namespace

void prof_begin(int kind, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.launches[kind]++;
  if (!g_prof.on) return;
  Prof::Pair p{g_prof.get(), g_prof.get(), kind};
  cudaEventRecord(p.a, st);
  g_prof.open_pairs.push_back(p);
}
void prof_end(int kind, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!g_prof.on || g_prof.open_pairs.empty()) return;
  Prof::Pair p = g_prof.open_pairs.back();
  g_prof.open_pairs.pop_back();
  cudaEventRecord(p.b, st);
  g_prof.done.push_back(p);
  (void)kind;
}

int check_mlp(const char* name, const psvo_mlp& net, int d_in, int d_out) {
  PSVO_REQUIRE(net.n_hidden >= 0 && net.n_hidden <= PSVO_MAX_HIDDEN, "%s: n_hidden=%d not in [0,%d]", name, net.n_hidden, PSVO_MAX_HIDDEN);
  PSVO_REQUIRE(net.d_in == d_in, "%s: d_in=%d, expected %d", name, net.d_in, d_in);
  PSVO_REQUIRE(net.d_out == d_out, "%s: d_out=%d, expected %d", name, net.d_out, d_out);
  for (int k = 0; k <= net.n_hidden; ++k) {
    PSVO_REQUIRE(net.W[k] != nullptr && net.b[k] != nullptr, "%s: layer %d has a null weight/bias pointer", name, k);
    if (k < net.n_hidden) PSVO_REQUIRE(net.width[k] >= 1 && net.width[k] <= 1024, "%s: hidden width %d out of range", name, net.width[k]);
  }
  return 0;
}

int check_dims(const psvo_dims* d) {
  PSVO_REQUIRE(d != nullptr, "dims is null");
  PSVO_REQUIRE(d->N >= 1 && d->B >= 1 && d->T >= 1, "N=%d B=%d T=%d must be positive", d->N, d->B, d->T);
  PSVO_REQUIRE(d->Dx >= 1 && d->Dx <= PSVO_MAX_DX, "Dx=%d not in [1,%d]", d->Dx, PSVO_MAX_DX);
  PSVO_REQUIRE(d->Dy >= 1 && d->Dy <= PSVO_MAX_DY, "Dy=%d not in [1,%d]", d->Dy, PSVO_MAX_DY);
  PSVO_REQUIRE((int64_t)d->B * d->T * d->N < (1LL << 40), "B*T*N too large");
  return 0;
}

struct FwdWs { float *sig, *mu2_all, *mu0, *muf0, *logWbar; unsigned char* tc_image; int* tc_sync; size_t bytes; };

FwdWs carve_fwd(const psvo_dims* d, void* ws) {
  FwdWs w;
  char* p = static_cast<char*>(ws);
  size_t off = 0;
  auto take = [&](size_t floats) { float* r = reinterpret_cast<float*>(p + off); off += align_up(floats * sizeof(float), 256); return r; };
  w.sig = take(SIG_SLOTS * 32);
  w.mu2_all = take((size_t)d->B * d->T * d->Dx);
  w.mu0 = take((size_t)d->B * d->Dx);
  w.muf0 = take((size_t)d->B * d->Dx);
  w.logWbar = take((size_t)d->B * d->N);
  w.tc_image = reinterpret_cast<unsigned char*>(take(32 * 1024));      // 128 KB: fp16 weight image of the tensor-core path
  w.tc_sync = reinterpret_cast<int*>(take(tc_filter_sync_bytes(d->B) / sizeof(float)));   // per-sequence flags of the persistent tensor-core filter
  w.bytes = off;
  return w;
}

SigmaSrc sigma_src(const psvo_dims* d, const psvo_params* p, bool ffbsi) {
  SigmaSrc s;
  for (int i = 0; i < SIG_SLOTS; ++i) { s.con[i] = nullptr; s.smin[i] = 1.0f; s.dim[i] = 0; }
  auto set = [&](int slot, const psvo_dist& dist, int dim) { s.con[slot] = dist.sigma_con; s.smin[slot] = dist.sigma_min; s.dim[slot] = dim; };
  set(SIG_Q0, p->q0, d->Dx);
  set(SIG_Q1, p->q1, d->Dx);
  if (d->flags & PSVO_F_TWO_Q) set(SIG_Q2, p->q2, d->Dx);
  if (!(d->flags & PSVO_F_BOOTSTRAP)) set(SIG_F, p->f, d->Dx);
  if (!(d->flags & PSVO_F_POISSON)) set(SIG_G, p->g, d->Dy);
  if (ffbsi) {
    set(SIG_BQI, p->bsim_q_init, d->Dx);
    set(SIG_Q1I, p->q1_inv, d->Dx);
    set(SIG_BQ2, p->bsim_q2, d->Dx);
  }
  return s;
}

int check_smc_params(const psvo_dims* d, const psvo_params* p) {
  PSVO_REQUIRE(p != nullptr, "params is null");
  const bool two_q = d->flags & PSVO_F_TWO_Q, boot = d->flags & PSVO_F_BOOTSTRAP;
  if (int e = check_mlp("q0", p->q0.net, d->Dx0, d->Dx)) return e;
  if (int e = check_mlp("q1", p->q1.net, d->Dx, d->Dx)) return e;
  if (int e = check_mlp("g", p->g.net, d->Dx, d->Dy)) return e;
  if (two_q) if (int e = check_mlp("q2", p->q2.net, d->Dobs, d->Dx)) return e;
  if (!boot) if (int e = check_mlp("f", p->f.net, d->Dx, d->Dx)) return e;
  PSVO_REQUIRE(p->q0.sigma_con && p->q1.sigma_con, "q0/q1 sigma_con is null");
  PSVO_REQUIRE(!two_q || p->q2.sigma_con, "q2 sigma_con is null");
  PSVO_REQUIRE(boot || p->f.sigma_con, "f sigma_con is null");
  PSVO_REQUIRE((d->flags & PSVO_F_POISSON) || p->g.sigma_con, "g sigma_con is null (Gaussian emission)");
  // f.log_prob(preprocessed_X0, X_0) feeds pre_x0 to a Dx->Dx network (SVO.py:91-92; the reference applies
  // X0_transformer first, SVO.py:328-329)
  PSVO_REQUIRE((boot && two_q) || d->Dx0 == d->Dx, "without (use_bootstrap and use_2_q) pre_x0 must have width Dx (X0_transformer output)");
  return 0;
}

}  // namespace psvo

using namespace psvo;

extern "C" {

PSVO_API int psvo_version(void) { return PSVO_VERSION; }

PSVO_API int psvo_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.on = on != 0;
  return 0;
}

PSVO_API int psvo_profile_read(double* ms_by_kind, int64_t* launches_by_kind) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (auto& p : g_prof.done) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess && ms_by_kind)
      ms_by_kind[p.kind] += ms;
    g_prof.pool.push_back(p.a); g_prof.pool.push_back(p.b);
  }
  g_prof.done.clear();
  for (int k = 0; k < PSVO_PROF_KINDS; ++k) {
    if (launches_by_kind) launches_by_kind[k] += g_prof.launches[k];
    g_prof.launches[k] = 0;
  }
  return 0;
}
PSVO_API const char* psvo_last_error(void) { return last_error().c_str(); }

PSVO_API int psvo_smc_forward_path(const psvo_dims* d, const psvo_params* p) {
  if (int e = check_dims(d)) return e;
  if (int e = check_smc_params(d, p)) return e;
  PropArgs pa;
  const bool boot = d->flags & PSVO_F_BOOTSTRAP;
  pa.q1 = p->q1.net; pa.f = boot ? p->q1.net : p->f.net; pa.g = p->g.net;
  pa.N = d->N; pa.B = d->B; pa.T = d->T; pa.Dx = d->Dx; pa.Dy = d->Dy; pa.t = 0; pa.flags = d->flags;
  pa.tc_image = nullptr;
  const PropPlan plan = plan_propagate(pa);
  if (tc_propagate_supported(pa)) {
    pa.tc_image = reinterpret_cast<const unsigned char*>(p);        // any non-null value: only the shape is examined here
    ResampleArgs ra; ra.N = d->N; ra.B = d->B; ra.T = d->T; ra.flags = d->flags;
    return tc_filter_persistent_supported(pa, ra) ? 3 : 2;
  }
  size_t psm = 0; int pblk = 0, csz = 0;
  if (persistent_filter_supported(pa, plan, &psm, &pblk)) return 1;
  return cluster_filter_supported(pa, &csz, &pblk, &psm) ? 4 : 0;
}

PSVO_API size_t psvo_smc_workspace_bytes(const psvo_dims* d, const psvo_params* p) {
  (void)p;
  if (d == nullptr) return 0;
  return carve_fwd(d, nullptr).bytes;
}

PSVO_API int psvo_smc_forward(const psvo_dims* d, const psvo_params* p, const float* obs, const float* pre_obs,
                     const float* pre_x0, const float* eps, const float* u, const int32_t* forced_anc, uint64_t seed,
                     float* Xs, float* X_anc, int32_t* anc, float* log_Ws, float* lse, float* logZ_b, void* ws,
                     size_t ws_bytes, psvo_stream_t stream) {
  if (int e = check_dims(d)) return e;
  if (int e = check_smc_params(d, p)) return e;
  PSVO_REQUIRE(obs && pre_x0 && Xs && X_anc && anc && log_Ws && lse && logZ_b && ws, "psvo_smc_forward: null buffer");
  const bool two_q = d->flags & PSVO_F_TWO_Q, boot = d->flags & PSVO_F_BOOTSTRAP, resample = d->flags & PSVO_F_RESAMPLE;
  PSVO_REQUIRE(!two_q || pre_obs, "psvo_smc_forward: pre_obs is null but use_2_q is set");
  const FwdWs w = carve_fwd(d, ws);
  PSVO_REQUIRE(ws_bytes >= w.bytes, "workspace too small: %zu < %zu", ws_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;

  if (int e = launch_sigma_block(sigma_src(d, p, false), w.sig, st)) return e;
  if (two_q) if (int e = launch_mlp_rows_forward(p->q2.net, pre_obs, (int64_t)d->B * d->T, w.mu2_all, st)) return e;
  if (int e = launch_mlp_rows_forward(p->q0.net, pre_x0, d->B, w.mu0, st)) return e;
  const psvo_dist& fd = boot ? p->q1 : p->f;
  if (!(boot && two_q)) if (int e = launch_mlp_rows_forward(fd.net, pre_x0, d->B, w.muf0, st)) return e;

  PropArgs pa;
  pa.q1 = p->q1.net; pa.f = fd.net; pa.g = p->g.net;
  pa.sig = w.sig; pa.obs = obs; pa.mu2_all = two_q ? w.mu2_all : nullptr; pa.mu0 = w.mu0; pa.muf0 = w.muf0;
  pa.eps = eps; pa.logWbar = resample ? nullptr : w.logWbar;
  pa.X_anc = X_anc; pa.Xs = Xs; pa.log_Ws = log_Ws;
  pa.seed = seed; pa.N = d->N; pa.B = d->B; pa.T = d->T; pa.Dx = d->Dx; pa.Dy = d->Dy; pa.t = 0;
  pa.seq_offset = d->seq_offset; pa.scratch_half = 0; pa.flags = d->flags; pa.tc_image = nullptr;
  const PropPlan plan = plan_propagate(pa);
  if (tc_propagate_supported(pa)) {
    PSVO_REQUIRE(tc_image_bytes(pa) <= 128 * 1024, "tensor-core weight image too large");
    if (int e = launch_tc_stage_image(pa, w.tc_image, st)) return e;
    pa.tc_image = w.tc_image;
  }

  ResampleArgs ra;
  ra.log_Ws = log_Ws; ra.Xs = Xs; ra.u = u; ra.forced_anc = forced_anc; ra.X_anc = X_anc; ra.anc = anc;
  ra.lse = lse; ra.logZ_b = logZ_b; ra.logWbar = w.logWbar; ra.seed = seed;
  ra.N = d->N; ra.B = d->B; ra.T = d->T; ra.Dx = d->Dx; ra.t = 0; ra.seq_offset = d->seq_offset; ra.flags = d->flags; ra.stream_id = 1u;

  {
    size_t psm = 0; int pblk = 0;
    if (persistent_filter_supported(pa, plan, &psm, &pblk)) return launch_filter_persistent(pa, ra, plan, st);
    int csz = 0;
    if (cluster_filter_supported(pa, &csz, &pblk, &psm)) { ra.skip_gather = 0; return launch_filter_cluster(pa, ra, st); }
  }
  // tensor-core path: the ancestor gather X_anc[t] = Xs[t][anc[t]] is folded into the propagate launch of t+1 (which
  // needs the parents anyway), so only the last step's resample launch moves particle states
  const bool fuse_gather = pa.tc_image != nullptr;
  if (tc_filter_persistent_supported(pa, ra)) {
    // the whole filter in ONE launch; only the last step's resampling (the one that gathers X_anc itself) is a launch of its own
    pa.X_anc_w = X_anc;
    if (int e = launch_filter_tc_persistent(pa, ra, w.tc_sync, st)) return e;
    ra.t = d->T - 1; ra.skip_gather = 0; ra.pdl = 0;
    return launch_resample(ra, st);
  }
  for (int t = 0; t < d->T; ++t) {
    pa.t = t; ra.t = t;
    pa.anc_prev = fuse_gather && t > 0 ? anc : nullptr; pa.X_anc_w = X_anc;
    ra.skip_gather = fuse_gather && t + 1 < d->T ? 1 : 0;
    // programmatic dependent launches inside the time loop (the first propagate launch follows ordinary kernels -- sigma
    // block, feature networks, weight image -- and keeps plain stream order)
    pa.pdl = t > 0 && pdl_enabled() ? 1 : 0;
    ra.pdl = pdl_enabled() ? 1 : 0;
    if (int e = launch_propagate(pa, plan, st)) return e;
    if (int e = launch_resample(ra, st)) return e;
  }
  return 0;
}

}  // extern "C"
