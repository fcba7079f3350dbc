"""IWAE: importance-weighted bound = the forward particle filter WITHOUT resampling, fed raw observations.

Flag preset over SVO (reference: src/SMC/IWAE.py:8-12):
    resample_particles = False   ancestors are the identity, log Wbar_t = log_W_t - LSE_n log_W_t stays differentiable,
                                 so d logZ_b / d log_alpha_t[n] is the FINAL softmax weight for every t (smc_backward.cu)
    smooth_obs = False           q2 / q0 read the raw observations
With T = 1 this is exactly the K-sample IWAE bound (tests/test_oracle.py::test_T1_no_resample_is_iwae_bound)."""
from .SVO import SVO


class IWAE(SVO):
    uses_encoder = False

    def __init__(self, model, FLAGS, name="log_ZSMC"):
        super().__init__(model, FLAGS, name=name)
        self.smooth_obs = False
        self.resample_particles = False
