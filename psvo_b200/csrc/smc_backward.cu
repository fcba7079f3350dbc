// smc_backward.cu -- reverse mode of the forward filter (placeholder until the kernels land).
#include "smc_kernels.h"
using namespace psvo;
extern "C" {
PSVO_API size_t psvo_smc_backward_workspace_bytes(const psvo_dims* d, const psvo_params* p) { (void)d; (void)p; return 256; }
PSVO_API int psvo_smc_backward(const psvo_dims* d, const psvo_params* p, const float* obs, const float* pre_obs,
                               const float* pre_x0, const float* eps, uint64_t seed, const float* Xs, const float* X_anc,
                               const int32_t* anc, const float* log_Ws, const float* lse, const float* g_logZ_b,
                               psvo_grads* grads, float* g_pre_obs, float* g_pre_x0, void* ws, size_t ws_bytes,
                               psvo_stream_t stream) {
  return fail("psvo_smc_backward: not implemented in this build");
}
}
