"""AESMC: auto-encoding SMC objective = the forward particle filter with multinomial resampling, fed RAW observations.

Flag preset over SVO (reference: src/SMC/AESMC.py:8-11):
    smooth_obs = False       q2 reads obs[:, t] (B, Dy) and q0 reads obs[:, 0]; no bi-LSTM encoder is built or run
    resample_particles = True (inherited)
Everything else -- kernels, reverse mode, keyword arguments of get_log_ZSMC -- is SVO's.  This is the class the
config-3 (linear-Gaussian, Kalman known answer) and config-5 (scaling sweep) workloads of bench.py run."""
from .SVO import SVO


class AESMC(SVO):
    uses_encoder = False

    def __init__(self, model, FLAGS, name="log_ZSMC"):
        super().__init__(model, FLAGS, name=name)
        self.smooth_obs = False
