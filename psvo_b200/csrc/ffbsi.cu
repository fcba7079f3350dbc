// ffbsi.cu -- FFBSi backward simulation (placeholder until the kernels land).
#include "smc_kernels.h"
using namespace psvo;
extern "C" {
PSVO_API size_t psvo_ffbsi_workspace_bytes(const psvo_dims* d, const psvo_params* p) { (void)d; (void)p; return 256; }
PSVO_API int psvo_ffbsi_forward(const psvo_dims* d, const psvo_params* p, const float* obs, const float* bs_feat,
                                const float* pre_x0, const float* Xs, const float* log_Ws, const float* lse,
                                const float* eps_bw, const float* u_pick, const float* u_res, const int32_t* forced_pick,
                                const int32_t* forced_anc2, uint64_t seed, float* bw_Xs, int32_t* pick, float* f_lp,
                                float* g_lp, float* Omega_or_bw_log_W, float* bw_X_anc, int32_t* anc2, float* logZ_b,
                                void* ws, size_t ws_bytes, psvo_stream_t stream) {
  return fail("psvo_ffbsi_forward: not implemented in this build");
}
}
