"""Likelihood plug-ins with the reference's interface (likelihoods/*.py), backed by the CUDA kernels.

Each class mirrors the reference class of the same name: ctor, ``get_metadata() -> (dim_y, dim_f, dim_p)``,
``ismulti()``, ``var_exp(Y, M, V, gh_points=None, Y_metadata=None)``, ``var_exp_derivatives(...)``, ``logpdf``,
``dlogp_df``, ``d2logp_df2`` (Categorical's derivative methods take the leading function index ``df``,
categorical.py:102,115; Gamma/Beta return 2-tuples, gamma.py:80-101), ``predictive(m, v)`` (Gauss-Hermite predictive
mean / variance, e.g. bernoulli.py:113-128) and ``log_predictive(Ytest, mu_F_star, v_F_star, num_samples)`` (Monte-Carlo
log predictive density, e.g. bernoulli.py:130-144).  All numbers come from ``hmogp_lik_var_exp`` / ``hmogp_lik_pointwise``
/ ``hmogp_lik_predictive`` (hetmogp_b200/csrc/lik_kernels.cu); numpy arrays in -> numpy arrays out, torch CUDA tensors in
-> tensors out.  ``log_predictive`` draws its function samples with ``numpy.random.normal`` in the reference's order and
shapes (so a seeded run reproduces the reference's estimate) and evaluates the N x num_samples log-densities on the GPU.
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

    # ---- prediction (svmogp.py:340-378 -> het_likelihood.py:133-164)
    gh_tensor = 10   # nodes per axis of the tensor grids: the reference's instance has cached the var_exp table by then

    def predictive(self, m, v, gh_points=None, Y_metadata=None):
        dy, F, P = self.get_metadata()
        (m, v), kind, dev = self._prep(m, v)
        N = int(m.reshape(-1, F).shape[0])
        m, v = m.reshape(N, F), v.reshape(N, F)
        mean, var = self._empty((N, P), dev), self._empty((N, P), dev)
        d = self._desc()
        check(lib.hmogp_lik_predictive(C.byref(d), N, ptr(m), ptr(v), ptr(mean), ptr(var), int(self.gh_tensor), kind, None))
        return mean, var

    _sample_layout = "NSD"   # F_samples (Ntest, num_samples, D) as bernoulli.py:131-137; "NDS" for the logpdf_sampling classes

    @staticmethod
    def _regroup(lp):
        return lp

    def log_predictive(self, Ytest, mu_F_star, v_F_star, num_samples):
        """-log(S) + logsumexp_s log p(y_n | f_ns), summed over n and -- as the reference does -- divided by num_samples."""
        mu_F_star, v_F_star = np.asarray(mu_F_star, dtype=np.float64), np.asarray(v_F_star, dtype=np.float64)
        Ytest = np.asarray(Ytest, dtype=np.float64)
        Ntest, D = mu_F_star.shape
        S = int(num_samples)
        Fs = np.empty((Ntest, S, D))
        for d in range(D):                                     # the reference's draw order: one (Ntest, S) block per function
            Fs[:, :, d] = np.random.normal(mu_F_star[:, d][:, None], np.sqrt(v_F_star[:, d][:, None]), size=(Ntest, S))
        if self._sample_layout == "NSD" and D > 1:
            Fs = Fs[:, :, :1]                                  # those classes evaluate logpdf(F_samples[:,:,0], Ytest)
        Fd = Fs.shape[2]
        yrep = np.repeat(Ytest.reshape(Ntest, 1), S, axis=1).reshape(-1)
        lp = np.asarray(self._pointwise(Fs.reshape(Ntest * S, Fd), yrep)[0]).reshape(Ntest, S)
        lp = self._regroup(lp)
        mx = lp.max(axis=-1, keepdims=True)
        mx = np.where(np.isfinite(mx), mx, 0.0)
        log_pred = -np.log(S) + (mx[:, 0] + np.log(np.exp(lp - mx).sum(axis=-1)))
        return (1.0 / S) * log_pred.sum()


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
    """likelihoods/hetgaussian.py:17-104."""
    name = "HetGaussian"
    spec = ("HetGaussian",)
    _sample_layout = "NDS"   # logpdf_sampling, hetgaussian.py:35-39,90-104

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

    _sample_layout = "NDS"   # logpdf_sampling, categorical.py:48-63,271-285

    @staticmethod
    def _regroup(lp):
        """categorical.py:57-62 stacks the per-sample probabilities sample-major (row s * N + n) and then reshapes the flat
        log-pmf to (N, S) in C order: the groups that log_predictive's logsumexp runs over mix rows and samples.  Kept as
        the reference computes it."""
        N, S = lp.shape
        return lp.T.reshape(-1).reshape(N, S)

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
    _sample_layout = "NDS"   # (the reference defines no log_predictive for Gamma / Beta; provided on both functions)

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
