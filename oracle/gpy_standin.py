"""Stand-in for the third-party symbols the reference hot path touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference (pmorenoz/HetMOGP) pins GPy==1.9.5, paramz (unpinned), climin==0.1a1,
matplotlib==2.2.2, scipy==1.1.0, numpy==1.14.3 (/root/reference/requirements.txt:2-6).
GPy / paramz / climin / matplotlib are absent from this image and cannot be
installed (no network), so the exact symbols the hot-path files import are
restated here from GPy 1.9.5's published behaviour.  Each is
pinned by a closed-form known-answer test in tests/test_oracle_standin.py.

Reference call sites served:
  GPy.kern.RBF.K                       util.py:161,178,197  svmogp.py:124
  GPy.kern.RBF.update_gradients_*      svmogp.py:116,140,142
  GPy.kern.RBF.gradients_X             svmogp.py:154,156
  GPy.kern.Coregionalize (.W, .B)      util.py:161,178  svmogp.py:141,143,156
  GPy.util.linalg.jitchol/dpotri/dpotrs  util.py:198,199  svmogp_inf.py:124,214
  GPy.util.choleskies.*                svmogp_inf.py:118,176,178,193,234
  GPy.util.misc.safe_exp/safe_square   every likelihoods/*.py
  GPy.likelihoods.Likelihood._gh_points  every likelihoods/*.py
  Posterior / LatentFunctionInference  svmogp_inf.py:6-7,48,181
"""
import sys
import types

import numpy as np
import scipy.linalg
from scipy.linalg import lapack


# --------------------------------------------------------------------------- misc
_lim_val = np.finfo(np.float64).max
_lim_val_exp = np.log(_lim_val)
_lim_val_square = np.sqrt(_lim_val)


def safe_exp(f):
    """GPy.util.misc.safe_exp: exp with the argument clamped at log(DBL_MAX)."""
    return np.exp(np.clip(f, -np.inf, _lim_val_exp))


def safe_square(f):
    """GPy.util.misc.safe_square: square with the argument clamped at sqrt(DBL_MAX)."""
    return np.clip(f, -np.inf, _lim_val_square) ** 2


# ------------------------------------------------------------------------- linalg
class LinAlgError(np.linalg.LinAlgError):
    pass


def jitchol(A, maxtries=5):
    """GPy.util.linalg.jitchol: dpotrf(lower); on failure retry with jitter
    mean(diag)*1e-6 * 10^k, k=0..maxtries-1.  No jitter when the first try succeeds."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            return scipy.linalg.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
        except Exception:
            jitter *= 10
        finally:
            num_tries += 1
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def dpotri(A, lower=1):
    """GPy.util.linalg.dpotri: inverse from a Cholesky factor, symmetrised."""
    A = np.asfortranarray(A)
    R, info = lapack.dpotri(A, lower=lower)
    if lower:
        R = np.tril(R) + np.tril(R, -1).T
    else:
        R = np.triu(R) + np.triu(R, 1).T
    return R, info


def dpotrs(A, B, lower=1):
    """GPy.util.linalg.dpotrs: solve A x = B given the Cholesky factor of A."""
    A = np.asfortranarray(A)
    return lapack.dpotrs(A, B, lower=lower)


# --------------------------------------------------------------------- choleskies
def flat_to_triang(flat_mat):
    """GPy.util.choleskies.flat_to_triang: (M(M+1)/2, D) -> (D, M, M), row-major
    over the lower triangle (numpy.tril_indices order)."""
    N, D = flat_mat.shape
    M = int((-1 + int(round(np.sqrt(8 * N + 1)))) // 2)
    ret = np.zeros((D, M, M))
    ii, jj = np.tril_indices(M)
    for d in range(D):
        ret[d, ii, jj] = flat_mat[:, d]
    return ret


def triang_to_flat(L):
    """GPy.util.choleskies.triang_to_flat: inverse of flat_to_triang (reads only
    the lower triangle)."""
    D, M, _ = L.shape
    N = M * (M + 1) // 2
    flat = np.empty((N, D))
    ii, jj = np.tril_indices(M)
    for d in range(D):
        flat[:, d] = L[d, ii, jj]
    return flat


# ---------------------------------------------------------------------- param shim
class _Param(np.ndarray):
    """Enough of paramz.Param for the reference's hot path: an ndarray carrying
    a .gradient (and .values)."""

    def __new__(cls, name, arr):
        obj = np.array(arr, dtype=float).view(cls)
        obj.name = name
        obj.gradient = np.zeros(obj.shape)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, "name", None)
        self.gradient = None

    @property
    def values(self):
        return np.asarray(self)


# -------------------------------------------------------------------------- kernels
class RBF(object):
    """GPy.kern.RBF (isotropic): K = variance * exp(-r^2/2), r = |x-x'|/lengthscale."""

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, name="rbf"):
        assert not ARD
        self.input_dim = input_dim
        self.variance = _Param("variance", np.atleast_1d(float(np.asarray(variance).ravel()[0])))
        if lengthscale is None:
            lengthscale = 1.0
        self.lengthscale = _Param("lengthscale", np.atleast_1d(float(np.asarray(lengthscale).ravel()[0])))
        self.name = name
        self.gradient = np.zeros(2)

    def copy(self):
        return RBF(self.input_dim, self.variance[0], self.lengthscale[0], name=self.name)

    def prod(self, other, name="mul"):
        return _Prod([self, other], name)

    def _unscaled_dist(self, X, X2=None):
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2.0 * X.dot(X.T) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices(r2.shape[0])] = 0.0
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)
        X1sq = np.sum(np.square(X), 1)
        X2sq = np.sum(np.square(X2), 1)
        r2 = -2.0 * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
        r2 = np.clip(r2, 0, np.inf)
        return np.sqrt(r2)

    def _scaled_dist(self, X, X2=None):
        return self._unscaled_dist(X, X2) / self.lengthscale[0]

    def K_of_r(self, r):
        return self.variance[0] * np.exp(-0.5 * r ** 2)

    def K(self, X, X2=None):
        return self.K_of_r(self._scaled_dist(X, X2))

    def Kdiag(self, X):
        return np.full(X.shape[0], self.variance[0])

    def update_gradients_full(self, dL_dK, X, X2=None):
        r = self._scaled_dist(X, X2)
        K = self.K_of_r(r)
        self.variance.gradient = np.atleast_1d(np.sum(K * dL_dK) / self.variance[0])
        dL_dr = (-r * K) * dL_dK
        self.lengthscale.gradient = np.atleast_1d(-np.sum(dL_dr * r) / self.lengthscale[0])
        self.gradient = np.array([self.variance.gradient[0], self.lengthscale.gradient[0]])

    def update_gradients_diag(self, dL_dKdiag, X):
        self.variance.gradient = np.atleast_1d(np.sum(dL_dKdiag))
        self.lengthscale.gradient = np.atleast_1d(0.0)
        self.gradient = np.array([self.variance.gradient[0], 0.0])

    def gradients_X(self, dL_dK, X, X2=None):
        r = self._scaled_dist(X, X2)
        with np.errstate(divide="ignore"):
            invdist = np.where(r != 0.0, 1.0 / np.where(r != 0.0, r, 1.0), 0.0)
        dL_dr = (-r * self.K_of_r(r)) * dL_dK
        tmp = invdist * dL_dr
        if X2 is None:
            tmp = tmp + tmp.T
            X2 = X
        grad = np.empty(X.shape, dtype=np.float64)
        for q in range(self.input_dim):
            np.sum(tmp * (X[:, q][:, None] - X2[:, q][None, :]), axis=1, out=grad[:, q])
        return grad / self.lengthscale[0] ** 2


class Coregionalize(object):
    """GPy.kern.Coregionalize: B = W W^T + diag(kappa)."""

    def __init__(self, input_dim, output_dim, rank=1, W=None, kappa=None, name="coregion"):
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.rank = rank
        if W is None:
            W = 0.5 * np.random.randn(output_dim, rank) / np.sqrt(rank)
        if kappa is None:
            kappa = 0.5 * np.ones(output_dim)
        self.W = _Param("W", np.asarray(W, dtype=float).reshape(output_dim, rank))
        self.kappa = _Param("kappa", np.asarray(kappa, dtype=float).reshape(output_dim))
        self.name = name

    @property
    def B(self):
        W = np.asarray(self.W)
        return W.dot(W.T) + np.diag(np.asarray(self.kappa))

    @property
    def gradient(self):
        return np.concatenate([np.asarray(self.W.gradient).ravel(), np.asarray(self.kappa.gradient).ravel()])

    @gradient.setter
    def gradient(self, g):
        g = np.asarray(g, dtype=float).ravel()
        nW = self.W.size
        self.W.gradient = g[:nW].reshape(self.W.shape)
        self.kappa.gradient = g[nW:].reshape(self.kappa.shape)


class _Prod(object):
    """Placeholder for the product/sum kernels util.ICM/LCM build and discard
    (util.py:116,122,142)."""

    def __init__(self, parts, name="mul"):
        self.parts = list(parts)
        self.name = name

    def __iadd__(self, other):
        self.parts.append(other)
        return self

    def __add__(self, other):
        return _Prod(self.parts + [other], self.name)


# ---------------------------------------------------------------------- likelihoods
class Identity(object):
    def transf(self, f):
        return f


class Likelihood(object):
    """GPy.likelihoods.Likelihood: only the constructor and the Gauss-Hermite
    table cache (first table built is kept; later T ignored -- SURVEY App. C-3)."""

    def __init__(self, gp_link, name):
        self.gp_link = gp_link
        self.name = name
        self.__gh_points = None

    def _gh_points(self, T=20):
        if self.__gh_points is None:
            self.__gh_points = np.polynomial.hermite.hermgauss(T)
        return self.__gh_points


class LatentFunctionInference(object):
    pass


class Posterior(object):
    """GPy Posterior: lazy container; the hot path only constructs it."""

    def __init__(self, woodbury_chol=None, woodbury_vector=None, K=None, mean=None, cov=None,
                 K_chol=None, woodbury_inv=None, prior_mean=0):
        self.mean = mean
        self.covariance = cov
        self._K = K
        self.prior_mean = prior_mean


def std_norm_pdf(x):
    return np.exp(-np.square(x) / 2) / np.sqrt(2 * np.pi)


def std_norm_cdf(x):
    from scipy.special import ndtr
    return ndtr(x)


# ------------------------------------------------------------------- module install
def install():
    """Register the stand-in under the module names the reference imports and add
    the two API shims newer numpy/scipy need (np.int, scipy.misc.logsumexp --
    categorical.py:81, util.py:234, bernoulli.py:10)."""
    if "GPy" in sys.modules and getattr(sys.modules["GPy"], "__hetmogp_standin__", False):
        return
    import scipy.special

    if not hasattr(np, "int"):
        np.int = int
    try:
        import scipy.misc as _misc
    except Exception:  # scipy.misc removed entirely
        _misc = types.ModuleType("scipy.misc")
        sys.modules["scipy.misc"] = _misc
        scipy.misc = _misc
    if not hasattr(_misc, "logsumexp"):
        _misc.logsumexp = scipy.special.logsumexp

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    linalg = mod("GPy.util.linalg", jitchol=jitchol, dpotri=dpotri, dpotrs=dpotrs)
    chol = mod("GPy.util.choleskies", flat_to_triang=flat_to_triang, triang_to_flat=triang_to_flat)
    misc = mod("GPy.util.misc", safe_exp=safe_exp, safe_square=safe_square)
    ug = mod("GPy.util.univariate_Gaussian", std_norm_pdf=std_norm_pdf, std_norm_cdf=std_norm_cdf)
    util = mod("GPy.util", linalg=linalg, choleskies=chol, misc=misc, univariate_Gaussian=ug)
    kern = mod("GPy.kern", RBF=RBF, Coregionalize=Coregionalize)
    links = mod("GPy.likelihoods.link_functions", Identity=Identity)
    liks = mod("GPy.likelihoods", Likelihood=Likelihood, link_functions=links)
    post = mod("GPy.inference.latent_function_inference.posterior", Posterior=Posterior)
    lfi = mod("GPy.inference.latent_function_inference", LatentFunctionInference=LatentFunctionInference,
              posterior=post)
    inf = mod("GPy.inference", latent_function_inference=lfi)
    gpy = mod("GPy", util=util, kern=kern, likelihoods=liks, inference=inf)
    gpy.__hetmogp_standin__ = True
    gpy.__path__ = []
    for name in ("matplotlib", "matplotlib.pyplot", "climin"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__hetmogp_standin__ = True
            sys.modules[name] = m
    if getattr(sys.modules["matplotlib"], "__hetmogp_standin__", False):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
