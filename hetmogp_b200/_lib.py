"""ctypes binding of libhetmogp_b200.so (C-ABI declared in include/hetmogp_b200.h).

The library is the product: there is NO Python/CPU fallback.  Importing this module fails loudly when the
shared object is missing (run ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C
hetmogp_b200/csrc``), and every compute entry point fails when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HMOGP_LIB") or os.path.join(_HERE, "lib", "libhetmogp_b200.so")   # HMOGP_LIB: development override

# constants mirrored from include/hetmogp_b200.h
MEM_HOST, MEM_DEVICE = 0, 1
PREC_FP64, PREC_FP32, PREC_TC = 0, 1, 2
WHAT_ELBO, WHAT_VE, WHAT_FULL = 0, 1, 2
ERR_ARG, ERR_CUDA, ERR_LINALG, ERR_UNSTABLE = 1, 2, 3, 4
LIK_KINDS = {"Gaussian": 0, "HetGaussian": 1, "Bernoulli": 2, "Poisson": 3, "Categorical": 4, "Gamma": 5, "Beta": 6,
             "Exponential": 7}
PRECISIONS = {"fp64": PREC_FP64, "fp32": PREC_FP32, "tc": PREC_TC}
MAX_Q, MAX_TASKS = 8, 16

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)
c_int32_p = C.POINTER(C.c_int32)


class LikDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("K", C.c_int32), ("sigma", C.c_double)]


class Config(C.Structure):
    _fields_ = [("M", C.c_int32), ("Q", C.c_int32), ("Xdim", C.c_int32), ("T", C.c_int32), ("precision", C.c_int32),
                ("device", C.c_int32), ("liks", C.POINTER(LikDesc))]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("Z", "m_u", "L_u", "rbf_var", "rbf_ls", "W", "kappa", "W_chain",
                                          "kappa_chain", "batch_scale")]


class Grads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("log_marginal", "VE", "KL", "dL_dmu_u", "dL_dL_u", "dL_dKmm", "d_rbf", "dW",
                                          "dkappa", "dZ")]


class Status(C.Structure):
    _fields_ = [("chol_fail", C.c_int32 * MAX_Q), ("jitter", C.c_double * MAX_Q), ("lu_singular", C.c_int32 * MAX_Q),
                ("n_negative_v", C.c_int64)]


class OptSegment(C.Structure):
    _fields_ = [("offset", C.c_int64), ("count", C.c_int64), ("param", C.c_void_p), ("grad", C.c_void_p),
                ("stride", C.c_int32), ("positive", C.c_int32), ("variational", C.c_int32), ("reserved", C.c_int32)]


class LinAlgError(np.linalg.LinAlgError):
    """jitchol gave up (GPy raises numpy.linalg.LinAlgError; reference util.py:198)."""


class HetMOGPError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "hetmogp_b200: %s not found. Build it with `make -C hetmogp_b200/csrc` (or __graft_entry__.build()). "
            "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    sig = {
        "hmogp_abi_version": (C.c_int, []),
        "hmogp_last_error": (C.c_char_p, []),
        "hmogp_device_count": (C.c_int, []),
        "hmogp_lik_dims": (C.c_int, [C.POINTER(LikDesc), c_int32_p, c_int32_p, c_int32_p]),
        "hmogp_generate_metadata": (C.c_int, [i32, C.POINTER(LikDesc), c_int64_p, c_int64_p, c_int64_p, c_int64_p,
                                              c_int64_p, c_int32_p, c_int32_p, c_int32_p]),
        "hmogp_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "hmogp_destroy": (None, [vp]),
        "hmogp_set_stream": (C.c_int, [vp, vp]),
        "hmogp_set_data": (C.c_int, [vp, i32, vp, vp, i64, i32]),
        "hmogp_set_rows": (C.c_int, [vp, c_int64_p, c_int64_p]),
        "hmogp_elbo_and_grads": (C.c_int, [vp, C.POINTER(Params), C.POINTER(Grads), i32, i32, C.POINTER(Status)]),
        "hmogp_stats_len": (i64, [vp]),
        "hmogp_stats_ptr": (vp, [vp]),
        "hmogp_step_local": (C.c_int, [vp, C.POINTER(Params), i32, i32, vp]),
        "hmogp_step_finish": (C.c_int, [vp, vp, C.POINTER(Grads), i32, i32, C.POINTER(Status)]),
        "hmogp_inference_host": (C.c_int, [C.POINTER(Config), C.POINTER(vp), C.POINTER(vp), c_int64_p,
                                           C.POINTER(Params), C.POINTER(Grads), i32, C.POINTER(Status)]),
        "hmogp_predict_f": (C.c_int, [vp, C.POINTER(Params), i32, i32, vp, i64, vp, vp]),
        "hmogp_get_rows": (C.c_int, [vp, i32, vp, vp, vp, vp, vp]),
        "hmogp_get_dL_dKmn": (C.c_int, [vp, i32, i32, vp, vp]),
        "hmogp_get_kuu": (C.c_int, [vp, vp, vp, vp]),
        "hmogp_lik_var_exp": (C.c_int, [C.POINTER(LikDesc), i64, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
        "hmogp_lik_pointwise": (C.c_int, [C.POINTER(LikDesc), i64, vp, vp, vp, vp, vp, i32, vp]),
        "hmogp_lik_predictive": (C.c_int, [C.POINTER(LikDesc), i64, vp, vp, vp, vp, i32, i32, vp]),
        "hmogp_flat_to_triang": (C.c_int, [vp, vp, i32, i32, i32, vp]),
        "hmogp_triang_to_flat": (C.c_int, [vp, vp, i32, i32, i32, vp]),
        "hmogp_hint_hyper_unchanged": (C.c_int, [vp, i32]),
        "hmogp_kuu_reuse_count": (C.c_int64, [vp]),
        "hmogp_enable_timing": (C.c_int, [vp, i32]),
        "hmogp_last_timing": (C.c_int, [vp, C.POINTER(C.c_float), c_int32_p]),
        "hmogp_tc_built": (C.c_int, []),
        "hmogp_opt_create": (C.c_int, [i32, C.POINTER(OptSegment), i32, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.POINTER(vp)]),
        "hmogp_opt_destroy": (None, [vp]),
        "hmogp_opt_size": (i64, [vp]),
        "hmogp_opt_state": (vp, [vp, i32]),
        "hmogp_opt_get_state": (C.c_int, [vp, i32, vp, vp]),
        "hmogp_opt_gather": (C.c_int, [vp, vp]),
        "hmogp_opt_lookahead": (C.c_int, [vp, i32, vp]),
        "hmogp_opt_update": (C.c_int, [vp, i32, i32, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()


def last_error():
    return lib.hmogp_last_error().decode("utf-8", "replace")


def check(rc):
    """Map C-ABI status codes to the exceptions the reference raises."""
    if rc == 0:
        return
    msg = last_error()
    if rc == ERR_LINALG:
        raise LinAlgError(msg)                       # GPy jitchol (util.py:198)
    if rc == ERR_UNSTABLE:
        raise ValueError(msg)                        # svmogp_inf.py:126-127
    if rc == ERR_ARG:
        raise ValueError(msg)
    raise HetMOGPError(msg)


def lik_desc(spec):
    """spec = ('Gaussian', sigma) | ('Categorical', K) | ('Bernoulli',) ..."""
    name = spec[0]
    d = LikDesc()
    d.kind = LIK_KINDS[name]
    d.K = int(spec[1]) if name == "Categorical" else 0
    d.sigma = float(spec[1]) if (name == "Gaussian" and len(spec) > 1 and spec[1] is not None) else 0.5
    return d


def ptr(a):
    """Raw pointer of a C-contiguous float64 numpy array or a torch tensor (host or CUDA); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need C-contiguous float64"
        return a.ctypes.data
    return a.data_ptr()  # torch tensor


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
