"""psvo_b200 -- the SMC objective of amoretti86/PSVO on NVIDIA B200 (sm_100a).

Host-side mirror of the reference's class interface (src/SMC, src/transformation, src/distribution, src/model.py)
over hand-written CUDA kernels reached through the C ABI in include/psvo.h.  There is no CPU path: every
operator raises if libpsvo_b200.so is missing or a tensor is not on a CUDA device."""
__version__ = "0.1.0"
