/* psvo.h -- C ABI of libpsvo_b200: the SMC-objective hot path of amoretti86/PSVO on NVIDIA B200.
 *
 * The reference has no FFI: the path sits behind a Python class API
 *   SVO / AESMC / IWAE / PSVO / PSVOwR (.get_log_ZSMC)      src/SMC/SVO.py:31-58, src/SMC/PSVO.py:21-50
 *   MLP_transformation.transform                            src/transformation/MLP.py:48-68
 *   tf_mvn / tf_poisson (.log_prob, .get_sigma)             src/distribution/mvn.py:80-107, poisson.py:33-45
 * Each entry point below names the reference function it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C structs, raw DEVICE pointers, sizes; no C++/torch types cross the boundary;
 *   - the caller owns every buffer (inputs, outputs, workspace); the library never allocates or frees
 *     device memory and keeps no pointer after return;
 *   - all work is enqueued on the given CUDA stream (a cudaStream_t passed as void*); calls return
 *     without synchronising;
 *   - return value: 0 ok; <0 argument/shape error (text via psvo_last_error()); >0 a cudaError_t;
 *   - fp32 arithmetic, int32 indices.  Tensors are sequence-major and contiguous:
 *       obs (B,T,Dy)  pre_obs (B,T,Dobs)  pre_x0 (B,Dx0)  Xs / X_anc (B,T,N,Dx)  log_Ws / anc (B,T,N)
 *     (the reference is particle-major (T,N,B,.) and transposes to (B,T,N,Dx) for log["Xs"],
 *      src/SMC/SVO.py:54);
 *   - MLP weights are Keras Dense kernels: row-major [in, out]; biases [out] (src/transformation/MLP.py:24-38).
 */
#ifndef PSVO_H_
#define PSVO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSVO_VERSION 200          /* 0.2.0 */
#define PSVO_MAX_HIDDEN 4         /* hidden layers per MLP_transformation */
#define PSVO_MAX_DX 16            /* latent dimension the fused kernels are instantiated for */
#define PSVO_MAX_DY 32            /* observation dimension the fused kernels are instantiated for */

/* flags (psvo_dims.flags) */
#define PSVO_F_BOOTSTRAP 1u       /* FLAGS.use_bootstrap : f aliases q1 (src/model.py:99-100,145-146) */
#define PSVO_F_TWO_Q 2u           /* FLAGS.use_2_q       : Gaussian fusion proposal (src/SMC/SVO.py:133-138) */
#define PSVO_F_RESAMPLE 4u        /* self.resample_particles (SVO.py:27; IWAE.py:11 clears it) */
#define PSVO_F_POISSON 8u         /* FLAGS.poisson_emission : tf_poisson emission (src/model.py:153-160) */
#define PSVO_F_WITH_RESAMPLING 16u/* PSVOwR variant of the backward simulation (src/SMC/PSVOwR.py) */
#define PSVO_F_SYSTEMATIC 32u     /* EXTRA, not in the reference (it only has multinomial, src/SMC/SVO.py:266-300): systematic
                                     resampling in the forward filter -- child n uses u_n = (n + u_0) / N with the single
                                     draw u_0 = u[b][t][0] (or its Philox value); same fixed-point CDF and search */

typedef void* psvo_stream_t;      /* cudaStream_t */

#if defined(__GNUC__)
#define PSVO_API __attribute__((visibility("default")))
#else
#define PSVO_API
#endif

/* MLP_transformation(Dhs, Dout): hidden ReLU Dense stack + linear mu head (src/transformation/MLP.py:24-54).
 * W[k]/b[k], k < n_hidden are the hidden layers, W[n_hidden]/b[n_hidden] the mu head. */
typedef struct psvo_mlp {
  int32_t d_in, d_out, n_hidden;
  int32_t width[PSVO_MAX_HIDDEN];
  const float* W[PSVO_MAX_HIDDEN + 1];
  const float* b[PSVO_MAX_HIDDEN + 1];
} psvo_mlp;

/* tf_mvn(transformation, sigma_init, sigma_min): sigma = max(softplus(sigma_con), sigma_min)
 * (src/distribution/mvn.py:80-90).  sigma_con == NULL means "unit scale" (tf_poisson, poisson.py:37). */
typedef struct psvo_dist {
  psvo_mlp net;
  const float* sigma_con;         /* [d_out] trainable pre-softplus stddev */
  float sigma_min;
  int32_t reserved;
} psvo_dist;

/* The SSM's distributions (src/model.py:113-160).  f is ignored when PSVO_F_BOOTSTRAP, q2 when
 * !PSVO_F_TWO_Q; the last three only matter to psvo_ffbsi_*. */
typedef struct psvo_params {
  psvo_dist q0, q1, q2, f, g;
  psvo_dist bsim_q_init, q1_inv, bsim_q2;
} psvo_params;

/* Gradient buffers, same topology as psvo_params; every non-NULL pointer is ACCUMULATED into (+=),
 * so the caller zero-fills.  With PSVO_F_BOOTSTRAP the f-gradients are added into q1's buffers,
 * which is what aliasing f_tran = q1_tran does under TF autodiff. */
typedef struct psvo_mlp_grad {
  float* W[PSVO_MAX_HIDDEN + 1];
  float* b[PSVO_MAX_HIDDEN + 1];
} psvo_mlp_grad;
typedef struct psvo_dist_grad {
  psvo_mlp_grad net;
  float* sigma_con;
} psvo_dist_grad;
typedef struct psvo_grads {
  psvo_dist_grad q0, q1, q2, f, g;
  psvo_dist_grad bsim_q_init, q1_inv, bsim_q2;
} psvo_grads;

typedef struct psvo_dims {
  int32_t N;          /* FLAGS.n_particles */
  int32_t M;          /* FLAGS.n_particles_for_BSim_proposal (ffbsi only) */
  int32_t B, T;       /* sequences on this device, time steps */
  int32_t Dx, Dy;
  int32_t Dobs;       /* width of pre_obs rows = q2 input */
  int32_t Dx0;        /* width of pre_x0 rows  = q0 input */
  int32_t Dfeat;      /* width of bs_feat rows = BSim_q_init / BSim_q2 input (ffbsi only) */
  int32_t seq_offset; /* global index of this device's first sequence: keys the in-kernel Philox so
                         that results do not depend on how the batch is sharded over GPUs */
  uint32_t flags;
  int32_t reserved;
} psvo_dims;

PSVO_API int psvo_version(void);
PSVO_API const char* psvo_last_error(void);   /* thread-local text of the last <0 return */

/* ---- forward particle filter ---------------------------------------------------------------------------
 * Replaces SVO.SMC + SVO.compute_log_ZSMC (src/SMC/SVO.py:60-180, 302-311) incl. sample_from_2_dist
 * (:182-197,228-232), resample_X / get_resample_idx (:243-300) and the q0/q2 feature MLPs.
 *   eps (B,T,N,Dx) and u (B,T,N): injected N(0,1) / U[0,1) draws (parity mode); NULL -> counter-based
 *   Philox4x32-10 keyed by seed (production mode; same streams as psvo_fill_normal / psvo_fill_uniform).
 *   forced_anc (B,T,N) or NULL: teacher-forced ancestor indices (parity tests).
 * Outputs: Xs = particles before resampling, X_anc = after (= log["Xs"]), anc = ancestor indices,
 *   log_Ws, lse (B,T) = logsumexp_n log_Ws, logZ_b (B) = sum_t lse  (log_ZSMC = mean_b logZ_b).
 *   A non-finite logZ_b marks a diverged sequence (the reference stops on it, src/trainer.py:221-223). */
PSVO_API size_t psvo_smc_workspace_bytes(const psvo_dims* d, const psvo_params* p);
PSVO_API int psvo_smc_forward(const psvo_dims* d, const psvo_params* p,
                     const float* obs, const float* pre_obs, const float* pre_x0,
                     const float* eps, const float* u, const int32_t* forced_anc, uint64_t seed,
                     float* Xs, float* X_anc, int32_t* anc, float* log_Ws, float* lse, float* logZ_b,
                     void* ws, size_t ws_bytes, psvo_stream_t stream);

/* ---- reverse mode of the forward filter ----------------------------------------------------------------
 * Replaces what optimizer.minimize(-log_ZSMC) derives through SVO.SMC (src/trainer.py:115-118): path-wise
 * gradients through the reparameterised samples and the ancestor gather; none into the resampling logits.
 * Consumes the forward outputs (recomputes MLP activations).  g_logZ_b (B) = d loss / d logZ_b.
 * g_Xs (B,T,N,Dx) / g_log_Ws (B,T,N) or NULL: additional incoming gradients w.r.t. the filter's particles and log
 * weights -- what psvo_ffbsi_backward returns, because the smoothing objectives read X_prevs and log_Ws
 * (src/SMC/PSVO.py:39-40).  g_pre_obs (B,T,Dobs) and g_pre_x0 (B,Dx0) are overwritten; grads are accumulated. */
PSVO_API size_t psvo_smc_backward_workspace_bytes(const psvo_dims* d, const psvo_params* p);
PSVO_API int psvo_smc_backward(const psvo_dims* d, const psvo_params* p,
                      const float* obs, const float* pre_obs, const float* pre_x0,
                      const float* eps, uint64_t seed,
                      const float* Xs, const float* X_anc, const int32_t* anc, const float* log_Ws,
                      const float* lse, const float* g_logZ_b, const float* g_Xs, const float* g_log_Ws,
                      psvo_grads* grads, float* g_pre_obs, float* g_pre_x0,
                      void* ws, size_t ws_bytes, psvo_stream_t stream);

/* ---- FFBSi backward simulation with an M-sample proposal ------------------------------------------------
 * Replaces PSVO.backward_simulation_w_proposal + PSVO.compute_log_ZSMC (src/SMC/PSVO.py:52-203) and, with
 * PSVO_F_WITH_RESAMPLING, the PSVOwR variants (src/SMC/PSVOwR.py:52-198).
 *   bs_feat (B,T,Dfeat): BS_preprocess_obs features (PSVO.py:205-216); pre_x0: the forward filter's
 *   cached preprocessed_X0 (PSVO.py:169); Xs, log_Ws, lse: outputs of psvo_smc_forward.
 *   eps_bw (B,T,M,N,Dx), u_pick (B,T,N), u_res (B,T,N; PSVOwR only) or NULL -> Philox(seed).
 *   forced_pick / forced_anc2 (B,T,N) or NULL: teacher forcing.
 * Outputs: bw_Xs (B,T,N,Dx), pick (B,T,N); PSVO: f_lp, g_lp, Omega (B,T,N); PSVOwR: bw_log_W in out_a,
 *   bw_X_anc (B,T,N,Dx) and anc2; omega_all (B,T,N,M) or NULL = the normalised log omega of every candidate
 *   (the logits of the 1-of-M pick; tests use it for the stage-wise bit-exactness check); logZ_b (B).
 *   Outputs of the other variant may be NULL. */
PSVO_API size_t psvo_ffbsi_workspace_bytes(const psvo_dims* d, const psvo_params* p);
PSVO_API int psvo_ffbsi_forward(const psvo_dims* d, const psvo_params* p,
                       const float* obs, const float* bs_feat, const float* pre_x0,
                       const float* Xs, const float* log_Ws, const float* lse,
                       const float* eps_bw, const float* u_pick, const float* u_res,
                       const int32_t* forced_pick, const int32_t* forced_anc2, uint64_t seed,
                       float* bw_Xs, int32_t* pick, float* f_lp, float* g_lp, float* Omega_or_bw_log_W,
                       float* bw_X_anc, int32_t* anc2, float* omega_all, float* logZ_b,
                       void* ws, size_t ws_bytes, psvo_stream_t stream);

/* ---- reverse mode of the backward simulation --------------------------------------------------------------
 * Replaces what optimizer.minimize(-log_ZSMC) (src/trainer.py:115-118) derives through
 * PSVO.backward_simulation_w_proposal + compute_log_ZSMC (src/SMC/PSVO.py:52-203) / the PSVOwR variants
 * (src/SMC/PSVOwR.py:52-198): path-wise gradients through the reparameterised candidates, the gather of the picked
 * candidate (and PSVOwR's second gather), the MLP log densities, the pairwise predictive term and the normalisation
 * over the M candidates; none into the 1-of-M pick / resampling logits.  Candidates are recomputed from eps_bw or
 * Philox(seed) and the stored picks.  Inputs after `seed` are outputs of psvo_ffbsi_forward.
 * Outputs: parameter gradients ACCUMULATED into grads (f-gradients into q1 under PSVO_F_BOOTSTRAP);
 *   g_bs_feat (B,T,Dfeat) or NULL and g_pre_x0 (B,Dx0) or NULL are overwritten;
 *   g_Xs (B,T,N,Dx) and g_log_Ws (B,T,N) are overwritten: feed them to psvo_smc_backward (with g_logZ_b = 0). */
PSVO_API size_t psvo_ffbsi_backward_workspace_bytes(const psvo_dims* d, const psvo_params* p);
PSVO_API int psvo_ffbsi_backward(const psvo_dims* d, const psvo_params* p,
                        const float* obs, const float* bs_feat, const float* pre_x0,
                        const float* Xs, const float* log_Ws, const float* lse,
                        const float* eps_bw, uint64_t seed,
                        const float* bw_Xs, const int32_t* pick, const float* f_lp, const float* g_lp,
                        const float* Omega_or_bw_log_W, const float* bw_X_anc, const int32_t* anc2,
                        const float* g_logZ_b,
                        psvo_grads* grads, float* g_bs_feat, float* g_pre_x0, float* g_Xs, float* g_log_Ws,
                        void* ws, size_t ws_bytes, psvo_stream_t stream);

/* ---- stand-alone operators (API parity + unit tests) ----------------------------------------------------
 * psvo_resample       SVO.get_resample_idx (SVO.py:266-300): logits (rows,K), u (rows,n) -> idx (rows,n)
 * psvo_gather         SVO.resample_X gather (SVO.py:254-257): src (rows,K,D), idx (rows,n) -> dst (rows,n,D)
 * psvo_mlp_forward    MLP_transformation.transform (MLP.py:48-54): x (rows,d_in) -> mu (rows,d_out)
 * psvo_mlp_backward   its reverse mode: g_mu (rows,d_out) -> g_x (rows,d_in, may be NULL), grads accumulated
 * psvo_get_sigma      tf_mvn.get_sigma (mvn.py:80-90): sigma_con (D) -> sigma (D)
 * psvo_mvn_log_prob   MultivariateNormalDiag(mu, sigma).log_prob(x) (mvn.py:63,104-107); sigma (D) or NULL = 1
 * psvo_logsumexp      tf.reduce_logsumexp over the last axis: x (rows,K) -> out (rows)
 * psvo_fill_normal / psvo_fill_uniform: materialise the in-kernel Philox streams: out[i], i < count, is draw
 *   number i of stream `stream_id` (0 eps, 1 u, 2 eps_bw, 3 u_pick, 4 u_res); `width` = draws per particle. */
PSVO_API int psvo_resample(const float* logits, const float* u, int64_t rows, int32_t K, int32_t n,
                  int32_t* idx, psvo_stream_t stream);
PSVO_API int psvo_gather(const float* src, const int32_t* idx, int64_t rows, int32_t K, int32_t n, int32_t D,
                float* dst, psvo_stream_t stream);
PSVO_API int psvo_mlp_forward(const psvo_mlp* net, const float* x, int64_t rows, float* mu, psvo_stream_t stream);
PSVO_API int psvo_mlp_backward(const psvo_mlp* net, const float* x, const float* g_mu, int64_t rows,
                      float* g_x, psvo_mlp_grad* grads, psvo_stream_t stream);
/* psvo_mlp_backward on sm_100a tensor cores (tcgen05, fp16x3 split operands, fp32 accumulation; weight gradients as
 * GEMMs over the rows with TMEM-resident accumulators).  One hidden layer of width 64/128/192/256, d_in <= 16,
 * d_out <= 32, rows >= 8192; psvo_mlp_backward_tc_supported() tells.  g_x is overwritten, grads are accumulated. */
PSVO_API int psvo_mlp_backward_tc_supported(const psvo_mlp* net, int64_t rows);
PSVO_API size_t psvo_mlp_backward_tc_workspace_bytes(const psvo_mlp* net);
PSVO_API int psvo_mlp_backward_tc(const psvo_mlp* net, const float* x, const float* g_mu, int64_t rows, float* g_x,
                         psvo_mlp_grad* grads, void* ws, size_t ws_bytes, psvo_stream_t stream);
PSVO_API int psvo_get_sigma(const float* sigma_con, float sigma_min, int32_t D, float* sigma, psvo_stream_t stream);
PSVO_API int psvo_mvn_log_prob(const float* x, const float* mu, const float* sigma, int64_t rows, int32_t D,
                      float* out, psvo_stream_t stream);
PSVO_API int psvo_logsumexp(const float* x, int64_t rows, int32_t K, float* out, psvo_stream_t stream);
PSVO_API int psvo_fill_normal(uint64_t seed, int32_t stream_id, int64_t first_particle, int64_t particles,
                     int32_t width, float* out, psvo_stream_t stream);
PSVO_API int psvo_fill_uniform(uint64_t seed, int32_t stream_id, int64_t first_particle, int64_t particles,
                      float* out, psvo_stream_t stream);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------
 * Launch counters are always on.  With profiling enabled every kernel launch is bracketed by CUDA events on
 * the launching stream; psvo_profile_read synchronises those events, ADDS per-kind device milliseconds and
 * launch counts into the two arrays (PSVO_PROF_KINDS entries each) and resets the internal state.
 * Kinds: 0 smc_propagate, 1 smc_resample, 2 mlp_rows (q0/q2 feature nets), 3 sigma, 4 smc_backward,
 *        5 ffbsi, 6 stand-alone operators. */
#define PSVO_PROF_KINDS 8
/* which kernels psvo_smc_forward would run for these dims / networks: 0 = per-step FP32 kernels (smc_propagate_kernel +
   smc_resample_kernel), 1 = persistent whole-filter kernel (smc_filter_persistent_kernel, one launch per pass),
   2 = per-step tcgen05 kernel (smc_propagate_tc_kernel) + resampling kernel (smc_resample_lean_kernel when every sequence gets a
   CTA of its own and N >= 2048, else smc_resample_kernel), 3 = the opt-in persistent tcgen05 whole-filter kernel
   (PSVO_TC_PERSIST=1: smc_filter_tc_persistent_kernel, resampling inside the launch), 4 = cluster whole-filter kernel
   (smc_filter_cluster_kernel: one launch per pass, one thread-block cluster per sequence; few sequences of many particles);
   <0 on invalid arguments */
PSVO_API int psvo_smc_forward_path(const psvo_dims* d, const psvo_params* p);
PSVO_API int psvo_profile_enable(int on);
PSVO_API int psvo_profile_read(double* ms_by_kind, int64_t* launches_by_kind);

#ifdef __cplusplus
}
#endif
#endif /* PSVO_H_ */
