"""ctypes binding of libpsvo_b200.so (include/psvo.h).  No CPU fallback: a missing library is an error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PSVO_LIB_PATH: development hook to load an experimental build of the same library (A/B kernel variants in one GPU session)
LIB_PATH = os.environ.get("PSVO_LIB_PATH") or os.path.join(_HERE, "lib", "libpsvo_b200.so")

MAX_HIDDEN = 4
F_BOOTSTRAP, F_TWO_Q, F_RESAMPLE, F_POISSON, F_WITH_RESAMPLING, F_SYSTEMATIC = 1, 2, 4, 8, 16, 32
f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class psvo_mlp(C.Structure):
    _fields_ = [("d_in", C.c_int32), ("d_out", C.c_int32), ("n_hidden", C.c_int32),
                ("width", C.c_int32 * MAX_HIDDEN),
                ("W", C.c_void_p * (MAX_HIDDEN + 1)), ("b", C.c_void_p * (MAX_HIDDEN + 1))]


class psvo_dist(C.Structure):
    _fields_ = [("net", psvo_mlp), ("sigma_con", C.c_void_p), ("sigma_min", C.c_float), ("reserved", C.c_int32)]


class psvo_params(C.Structure):
    _fields_ = [(k, psvo_dist) for k in ("q0", "q1", "q2", "f", "g", "bsim_q_init", "q1_inv", "bsim_q2")]


class psvo_mlp_grad(C.Structure):
    _fields_ = [("W", C.c_void_p * (MAX_HIDDEN + 1)), ("b", C.c_void_p * (MAX_HIDDEN + 1))]


class psvo_dist_grad(C.Structure):
    _fields_ = [("net", psvo_mlp_grad), ("sigma_con", C.c_void_p)]


class psvo_grads(C.Structure):
    _fields_ = [(k, psvo_dist_grad) for k in ("q0", "q1", "q2", "f", "g", "bsim_q_init", "q1_inv", "bsim_q2")]


class psvo_dims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("N", "M", "B", "T", "Dx", "Dy", "Dobs", "Dx0", "Dfeat", "seq_offset")] + \
               [("flags", C.c_uint32), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol include/psvo.h declares
_vp, _i32, _i64, _u64, _sz, _f = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t, C.c_float
_dp, _pp = C.POINTER(psvo_dims), C.POINTER(psvo_params)
SIGNATURES = {
    "psvo_version": (C.c_int, []),
    "psvo_last_error": (C.c_char_p, []),
    "psvo_smc_workspace_bytes": (_sz, [_dp, _pp]),
    "psvo_smc_forward_path": (C.c_int, [_dp, _pp]),
    "psvo_smc_forward": (C.c_int, [_dp, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _u64,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "psvo_smc_backward_workspace_bytes": (_sz, [_dp, _pp]),
    "psvo_smc_backward": (C.c_int, [_dp, _pp, _vp, _vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                    C.POINTER(psvo_grads), _vp, _vp, _vp, _sz, _vp]),
    "psvo_ffbsi_backward_workspace_bytes": (_sz, [_dp, _pp]),
    "psvo_ffbsi_backward": (C.c_int, [_dp, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                      C.POINTER(psvo_grads), _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "psvo_ffbsi_workspace_bytes": (_sz, [_dp, _pp]),
    "psvo_ffbsi_forward": (C.c_int, [_dp, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _u64,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "psvo_resample": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "psvo_gather": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "psvo_mlp_forward": (C.c_int, [C.POINTER(psvo_mlp), _vp, _i64, _vp, _vp]),
    "psvo_mlp_backward": (C.c_int, [C.POINTER(psvo_mlp), _vp, _vp, _i64, _vp, C.POINTER(psvo_mlp_grad), _vp]),
    "psvo_mlp_backward_tc_supported": (C.c_int, [C.POINTER(psvo_mlp), _i64]),
    "psvo_mlp_backward_tc_workspace_bytes": (_sz, [C.POINTER(psvo_mlp)]),
    "psvo_mlp_backward_tc": (C.c_int, [C.POINTER(psvo_mlp), _vp, _vp, _i64, _vp, C.POINTER(psvo_mlp_grad), _vp, _sz, _vp]),
    "psvo_get_sigma": (C.c_int, [_vp, _f, _i32, _vp, _vp]),
    "psvo_mvn_log_prob": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "psvo_logsumexp": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "psvo_fill_normal": (C.c_int, [_u64, _i32, _i64, _i64, _i32, _vp, _vp]),
    "psvo_fill_uniform": (C.c_int, [_u64, _i32, _i64, _i64, _vp, _vp]),
    "psvo_profile_enable": (C.c_int, [C.c_int]),
    "psvo_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libpsvo_b200.so not found at %s -- build it with `python -m psvo_b200.build` "
                               "(needs nvcc); psvo_b200 has no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class PsvoError(RuntimeError):
    pass


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise ValueError("%s: %s" % (what, load().psvo_last_error().decode()))
    raise PsvoError("%s: CUDA error %d (%s)" % (what, rc, load().psvo_last_error().decode()))
