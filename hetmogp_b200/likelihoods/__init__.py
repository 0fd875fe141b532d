"""Likelihood plug-ins with the reference's interface (likelihoods/*.py), backed by the CUDA kernels.

Each class mirrors the reference class of the same name: ctor, ``get_metadata() -> (dim_y, dim_f, dim_p)``,
``ismulti()``, ``var_exp(Y, M, V, gh_points=None, Y_metadata=None)``, ``var_exp_derivatives(...)``, ``logpdf``,
``dlogp_df``, ``d2logp_df2`` (Categorical's derivative methods take the leading function index ``df``,
categorical.py:102,115; Gamma/Beta return 2-tuples, gamma.py:80-101).  Prediction/sampling methods are out of
scope.  All numbers come from ``hmogp_lik_var_exp`` / ``hmogp_lik_pointwise``
(hetmogp_b200/csrc/lik_kernels.cu); numpy arrays in -> numpy arrays out, torch CUDA tensors in -> tensors out.
"""
import ctypes as C

import numpy as np

from .. import _lib
from .._lib import lib, check, ptr, f64


class _Likelihood(object):
    name = None
    spec = None
    precision = "fp64"   # arithmetic of the quadrature kernel for stand-alone calls ("fp64" | "fp32")

    def _desc(self):
        return _lib.lik_desc(self.spec)

    def get_metadata(self):
        d = self._desc()
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.hmogp_lik_dims(C.byref(d), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def ismulti(self):
        return False

    # ---- CUDA calls
    def _prep(self, *arrs):
        if any(hasattr(a, "data_ptr") for a in arrs):
            import torch
            dev = [a for a in arrs if hasattr(a, "data_ptr")][0].device
            out = [torch.as_tensor(a, dtype=torch.float64, device=dev).contiguous() for a in arrs]
            return out, _lib.MEM_DEVICE, dev
        return [f64(a) for a in arrs], _lib.MEM_HOST, None

    def _empty(self, shape, dev):
        if dev is None:
            return np.empty(shape)
        import torch
        return torch.empty(shape, dtype=torch.float64, device=dev)

    def _var_exp_all(self, Y, M, V, want=(True, True, True)):
        F = self.get_metadata()[1]
        (Y, M, V), kind, dev = self._prep(Y, M, V)
        N = int(Y.reshape(-1).shape[0])
        Y, M, V = Y.reshape(N), M.reshape(N, F), V.reshape(N, F)
        ve = self._empty((N, 1), dev) if want[0] else None
        dm = self._empty((N, F), dev) if want[1] else None
        dv = self._empty((N, F), dev) if want[2] else None
        d = self._desc()
        check(lib.hmogp_lik_var_exp(C.byref(d), N, ptr(Y), ptr(M), ptr(V), ptr(ve), ptr(dm), ptr(dv),
                                    _lib.PRECISIONS[self.precision], kind, None))
        return ve, dm, dv

    def _pointwise(self, F_, y):
        F = self.get_metadata()[1]
        (F_, y), kind, dev = self._prep(F_, y)
        N = int(y.reshape(-1).shape[0])
        F_, y = F_.reshape(N, F), y.reshape(N)
        lp, d1, d2 = self._empty((N,), dev), self._empty((N, F), dev), self._empty((N, F), dev)
        d = self._desc()
        check(lib.hmogp_lik_pointwise(C.byref(d), N, ptr(F_), ptr(y), ptr(lp), ptr(d1), ptr(d2), kind, None))
        return lp, d1, d2

    # ---- reference interface
    def var_exp(self, Y, M, V, gh_points=None, Y_metadata=None):
        return self._var_exp_all(Y, M, V, (True, False, False))[0]

    def var_exp_derivatives(self, Y, M, V, gh_points=None, Y_metadata=None):
        _, dm, dv = self._var_exp_all(Y, M, V, (False, True, True))
        return dm, dv

    def logpdf(self, F, y, Y_metadata=None):
        return self._pointwise(F, y)[0]


class _WithDerivs(_Likelihood):
    def dlogp_df(self, f, y, Y_metadata=None):
        return self._pointwise(f, y)[1]

    def d2logp_df2(self, f, y, Y_metadata=None):
        return self._pointwise(f, y)[2]


class Gaussian(_Likelihood):
    """likelihoods/gaussian.py:17-62."""
    name = "Gaussian"

    def __init__(self, sigma=None, gp_link=None):
        self.sigma = 0.5 if sigma is None else sigma   # gaussian.py:22
        self.spec = ("Gaussian", self.sigma)


class HetGaussian(_Likelihood):
    """likelihoods/hetgaussian.py:17-73."""
    name = "HetGaussian"
    spec = ("HetGaussian",)

    def __init__(self, gp_link=None):
        pass


class Bernoulli(_WithDerivs):
    """likelihoods/bernoulli.py:19-111."""
    name = "Bernoulli"
    spec = ("Bernoulli",)

    def __init__(self, gp_link=None):
        pass


class Poisson(_WithDerivs):
    """likelihoods/poisson.py."""
    name = "Poisson"
    spec = ("Poisson",)

    def __init__(self, gp_link=None):
        pass


class Exponential(_WithDerivs):
    """likelihoods/exponential.py:28-99."""
    name = "Exponential"
    spec = ("Exponential",)

    def __init__(self, gp_link=None):
        pass


class Categorical(_Likelihood):
    """likelihoods/categorical.py:22-222."""
    name = "Categorical"

    def __init__(self, K, gp_link=None):
        self.K = K
        self.spec = ("Categorical", K)

    def ismulti(self):
        return True

    def dlogp_df(self, df, F, y, Y_metadata=None):
        return self._pointwise(F, y)[1][:, df:df + 1]

    def d2logp_df2(self, df, F, y, Y_metadata=None):
        return self._pointwise(F, y)[2][:, df:df + 1]


class _TwoParam(_Likelihood):
    def dlogp_df(self, F, y, Y_metadata=None):
        d1 = self._pointwise(F, y)[1]
        return d1[:, 0:1], d1[:, 1:2]

    def d2logp_df2(self, F, y, Y_metadata=None):
        d2 = self._pointwise(F, y)[2]
        return d2[:, 0:1], d2[:, 1:2]


class Gamma(_TwoParam):
    """likelihoods/gamma.py:34-194."""
    name = "Gamma"
    spec = ("Gamma",)

    def __init__(self, gp_link=None):
        pass


class Beta(_TwoParam):
    """likelihoods/beta.py:29-197."""
    name = "Beta"
    spec = ("Beta",)

    def __init__(self, gp_link=None):
        pass


_BY_NAME = {c.name: c for c in (Gaussian, HetGaussian, Bernoulli, Poisson, Exponential, Categorical, Gamma, Beta)}


def from_spec(spec):
    name = spec[0]
    if name == "Gaussian":
        return Gaussian(spec[1] if len(spec) > 1 else None)
    if name == "Categorical":
        return Categorical(spec[1])
    return _BY_NAME[name]()
