// smc_kernels.h -- argument blocks and launchers shared by the CUDA translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "psvo_common.cuh"

namespace psvo {

// slots of the sigma block (32 floats each) computed by launch_sigma_block
enum { SIG_Q0 = 0, SIG_Q1, SIG_Q2, SIG_F, SIG_G, SIG_BQI, SIG_Q1I, SIG_BQ2, SIG_SLOTS };

struct SigmaSrc {
  const float* con[SIG_SLOTS];
  float smin[SIG_SLOTS];
  int dim[SIG_SLOTS];
};
int launch_sigma_block(const SigmaSrc& src, float* sig, cudaStream_t st);

struct PropArgs {
  psvo_mlp q1, f, g;
  const float* sig;        // sigma block
  const float* obs;        // (B,T,Dy)
  const float* mu2_all;    // (B,T,Dx) q2 means, or NULL when !two_q
  const float* mu0;        // (B,Dx)   q0 mean
  const float* muf0;       // (B,Dx)   f-dist mean of pre_x0 (only when !(bootstrap && two_q))
  const float* eps;        // (B,T,N,Dx) or NULL -> Philox
  const float* logWbar;    // (B,N) normalised log weights of t-1 (no-resample mode) or NULL -> -log N
  const float* X_anc;      // (B,T,N,Dx) read at t-1
  float* Xs;               // (B,T,N,Dx) written at t
  float* log_Ws;           // (B,T,N)    written at t
  const unsigned char* tc_image;   // staged fp16 weight image for the tensor-core path (launch_tc_stage_image) or NULL
  // fused ancestor gather (tensor-core path): when anc_prev != NULL the parent of particle n is read as Xs[t-1][anc_prev[t-1][n]]
  // and copied to X_anc_w[t-1][n]; the resample launch of t-1 then skipped its own gather (ResampleArgs::skip_gather)
  const int* anc_prev = nullptr;   // (B,T,N)
  float* X_anc_w = nullptr;        // (B,T,N,Dx)
  unsigned long long seed;
  int N, B, T, Dx, Dy, t, seq_offset, scratch_half;
  unsigned flags;
  int pdl = 0;             // 1: launch with programmatic stream serialisation (the kernel's prologue overlaps the previous kernel's tail)
};
struct PropPlan { int block, grid, ppt, scratch_half; size_t smem; };
PropPlan plan_propagate(const PropArgs& a);
int launch_propagate(PropArgs a, const PropPlan& pl, cudaStream_t st);

struct ResampleArgs {
  const float* log_Ws;     // (B,T,N)
  const float* Xs;         // (B,T,N,Dx)
  const float* u;          // (B,T,N) or NULL -> Philox
  const int* forced_anc;   // (B,T,N) or NULL
  float* X_anc;            // (B,T,N,Dx)
  int* anc;                // (B,T,N)
  float* lse;              // (B,T)
  float* logZ_b;           // (B)
  float* logWbar;          // (B,N) written when !resample
  unsigned long long seed;
  int N, B, T, Dx, t, seq_offset;
  unsigned flags;
  unsigned stream_id;      // Philox stream of the uniforms (1 forward filter, 4 PSVOwR second resampling)
  int pdl = 0;             // 1: programmatic dependent launch (see PropArgs::pdl)
  int skip_gather = 0;     // 1: write anc only; X_anc[t] is produced by the propagate launch of t+1 (PropArgs::anc_prev)
};
int launch_resample(const ResampleArgs& a, cudaStream_t st);

// persistent whole-filter kernel (one launch per pass, one CTA per sequence): small N / few sequences
bool persistent_env();     // PSVO_PERSISTENT != 0: whole-filter kernels allowed
bool persistent_filter_supported(const PropArgs& a, const PropPlan& pl, size_t* smem_out, int* block_out);
int launch_filter_persistent(PropArgs pa, const ResampleArgs& ra, const PropPlan& pl, cudaStream_t st);
// cluster whole-filter kernel (one launch per pass, one thread-block cluster per sequence): few sequences of many particles
bool cluster_filter_supported(const PropArgs& a, int* csize_out, int* block_out, size_t* smem_out);
int launch_filter_cluster(PropArgs pa, const ResampleArgs& ra, cudaStream_t st);

// generic row MLP (mlp_rows.cu)
int launch_mlp_rows_forward(const psvo_mlp& net, const float* x, int64_t rows, float* mu, cudaStream_t st);
int launch_mlp_rows_backward(const psvo_mlp& net, const float* x, const float* g_mu, int64_t rows,
                             float* g_x, const psvo_mlp_grad* grads, cudaStream_t st);

// per-particle Dx -> Dx MLP over (B,N) rows read with a per-sequence stride (ffbsi.cu)
int launch_particle_mlp_dx(const psvo_mlp& net, const float* in, long long seq_stride, int B, int N, int Dx, float* out,
                           cudaStream_t st);

// tensor-core (tcgen05) pairwise predictive term of the FFBSi backward simulation (ffbsi_pair_tc.cu)
struct PairTcArgs {
  const float* sig; const float* mu_f; const float* mu1i; const float* m2_all; const float* mu_init_all;
  const float* log_Ws; const float* lse; const float* eps_bw;
  unsigned char* bimg;     // B image workspace, ffbsi_pair_tc_image_bytes()
  int prep_all = 0;        // prep kernel only: 1 = the B images of ALL steps t = 1..T-1 in one launch (batch 1: bimg holds T slices,
                           // mu_f points at the (T,N,Dx) means of all steps)
  int prep_done = 0;       // launch_ffbsi_pair_tc: the B image of this step is already in bimg
  float* candP;            // (B, N, M) out: LSE_j( -|x~/s - mu_j/s|^2 / 2 + log Wbar_j )
  unsigned long long seed;
  int N, M, B, T, Dx, t, seq_offset;
  unsigned flags;
};
bool ffbsi_pair_tc_supported(int N, int M, int B, int Dx);
size_t ffbsi_pair_tc_image_bytes(int B, int N);
int launch_ffbsi_pair_tc(const PairTcArgs& a, cudaStream_t st);
int launch_ffbsi_pair_prep_all(PairTcArgs a, cudaStream_t st);     // a.bimg: T slices, a.mu_f: (T,N,Dx)

// tensor-core (tcgen05) propagate step (smc_propagate_tc.cu)
bool tc_propagate_supported(const PropArgs& a);
size_t tc_image_bytes(const PropArgs& a);
int launch_tc_stage_image(const PropArgs& a, unsigned char* image, cudaStream_t st);
int launch_propagate_tc(const PropArgs& a, cudaStream_t st);

// persistent tensor-core filter: ALL time steps of the forward filter in one cooperative launch; the resampling of steps
// 0..T-2 runs inside it (one worker warp per CTA, per-sequence flags in global memory), the caller launches the ordinary
// resampling kernel for step T-1 (the only one that gathers X_anc itself)
struct TcPersistArgs {
  ResampleArgs ra;
  int* cnt;                // (B) tiles stored per sequence, all steps; zeroed by the launcher
  int* ready;              // (B) resampled steps per sequence
  int* err;                // watchdog flag
  int* next;               // tile counter: tiles are handed out in (t, b, tile) order
  unsigned rs_smem_off;    // byte offset of the worker's region in dynamic shared memory
};
bool tc_filter_persistent_supported(const PropArgs& a, const ResampleArgs& ra);
size_t tc_filter_sync_bytes(int B);
int launch_filter_tc_persistent(PropArgs a, const ResampleArgs& ra, int* sync, cudaStream_t st);

// tensor-core (tcgen05) MLP f and MLP g of every FFBSi candidate (smc_propagate_tc.cu); weight image from
// launch_tc_stage_image with PropArgs.q1 = the f network
struct CandTcArgs {
  const unsigned char* tc_image; const float* sig; const float* obs;
  const float* mu1i; const float* m2_all; const float* mu_init_all; const float* X_next; const float* eps_bw;
  float* candF; float* candG;        // (B, N, M) out: f_lp (0 at t = T-1) and g_lp of every candidate
  float* candQ = nullptr;            // (B, N, M) out (optional): proposal log-density q(x~) of every candidate
  unsigned long long seed;
  int N, M, B, T, Dx, Dy, H, t, seq_offset;
  unsigned flags;
};
bool ffbsi_cand_tc_supported(const psvo_mlp& f, const psvo_mlp& g, int Dx, int Dy, long long rows);
int launch_ffbsi_cand_tc(const CandTcArgs& a, cudaStream_t st);

// tensor-core (tcgen05) forward of MLP q1 (on xq rows, optional) and MLP g (on xg rows) for the reverse pass
// (smc_propagate_tc.cu); weight image from launch_tc_stage_image
struct Fwd2TcArgs {
  const unsigned char* tc_image;
  const float* xq; const float* xg; long long x_seq_stride;     // row (b, n) at x + b*stride + n*Dx; xq may be NULL
  float* muq; float* mug;                                       // (B*N, Dx) and (B*N, Dy)
  int B, N, Dx, Dy, H;
  int prof_kind = PK_BWD;                                       // which profile counter the launch goes to
};
int launch_tc_fwd2(const Fwd2TcArgs& a, cudaStream_t st);     // xg == NULL: only the network in the "q" slot (xq -> muq)
bool tc_pipeline_supported(const psvo_mlp& q, const psvo_mlp& g, int Dx, int Dy);

// tensor-core (tcgen05) reverse mode of a one-hidden-layer MLP over particle rows (mlp_bwd_tc.cu)
bool mlp_bwd_tc_supported(const psvo_mlp& net, long long rows);
size_t mlp_bwd_tc_image_bytes(const psvo_mlp& net);
int launch_mlp_bwd_tc_stage(const psvo_mlp& net, unsigned char* image, cudaStream_t st);
int launch_mlp_bwd_tc(const psvo_mlp& net, const unsigned char* image, const float* x, long long x_seq_stride, const float* g,
                      int B, int N, float* gx, const int* scatter, long long scatter_seq_stride, const psvo_mlp_grad& grads,
                      unsigned int* gmax_bits, cudaStream_t st, bool gmax_ready = false);   // gmax_ready: *gmax_bits already holds max |g|

// tensor-core (tcgen05) one-hidden-layer MLP over particle rows (tc_mlp.cu)
bool tc_mlp_supported(const psvo_mlp& net);
int launch_particle_mlp_tc(const psvo_mlp& net, const float* in, long long seq_stride, int B, int N, float* out,
                           cudaStream_t st);

}  // namespace psvo
