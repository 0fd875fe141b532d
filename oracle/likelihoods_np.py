"""numpy restatement of the reference likelihood plug-ins (hot-path methods only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Travels to the GPU box (it
does not read /root/reference).  Each class follows the reference file cited in
its docstring; results agree with the verbatim reference to ~1e-13
(tests/test_oracle_golden.py, fixtures in tests/golden/).

Reference quirks reproduced: Categorical dlogp_df is the
constant 1[y=d+1]-1 (C-1); Gamma/Beta expectations carry an extra 1/pi (C-2);
Gauss-Hermite order is 20 for 1-D likelihoods and 10 per axis for tensor grids
(C-3).

Third-party arithmetic restated: scipy.stats.multinomial.logpmf(x=onehot, n=1, p)
== log p_y (scipy>=1.2 semantics; scipy 1.1.0, the reference's pin, recomputes
p_K = 1 - sum_{k<K} p_k first, a <=1e-16 relative difference).
"""
import numpy as np
from scipy.special import gammaln, psi, zeta, betaln

_LIM_EXP = np.log(np.finfo(np.float64).max)
_LIM_SQ = np.sqrt(np.finfo(np.float64).max)


def safe_exp(f):
    return np.exp(np.minimum(f, _LIM_EXP))


def safe_square(f):
    return np.minimum(f, _LIM_SQ) ** 2


def gh_points(T):
    """GPy Likelihood._gh_points -> numpy.polynomial.hermite.hermgauss(T)."""
    return np.polynomial.hermite.hermgauss(T)


def _grid(M, V, gh_f):
    """Tensor Gauss-Hermite grid (categorical.py:137-157): list of D arrays that
    broadcast to (N, T, ..., T); axis d+1 carries function d (C order)."""
    N, D = M.shape
    T = gh_f.shape[0]
    F = []
    for d in range(D):
        shp = [1] * (D + 1)
        shp[d + 1] = T
        mv = [N] + [1] * D
        F.append(gh_f.reshape(shp) * np.sqrt(2 * V[:, d].reshape(mv)) + M[:, d].reshape(mv))
    return F


def _contract(A, gh_w, D, prescaled):
    """logp.dot(gh_w)/sqrt(pi) repeated D times (categorical.py:165-168).  With
    prescaled=True gh_w was already divided by sqrt(pi) -- the Gamma/Beta double
    normalisation (gamma.py:110,139-141)."""
    w = gh_w / np.sqrt(np.pi) if prescaled else gh_w
    out = A
    for _ in range(D):
        out = out.dot(w) / np.sqrt(np.pi)
    return out


class Gaussian(object):
    """likelihoods/gaussian.py:17-62 (analytic)."""
    name = "Gaussian"
    dims = (1, 1, 1)

    def __init__(self, sigma=None):
        self.sigma = 0.5 if sigma is None else sigma

    def get_metadata(self):
        return self.dims

    def var_exp(self, Y, M, V):
        lik_v = np.square(self.sigma)
        m, v, y = M.reshape(-1, 1), V.reshape(-1, 1), Y.reshape(-1, 1)
        return -0.5 * np.log(2 * np.pi) - 0.5 * np.log(lik_v) \
            - 0.5 * (np.square(y) + np.square(m) + v - (2 * m * y)) / lik_v

    def var_exp_derivatives(self, Y, M, V):
        lik_v = np.square(self.sigma)
        m, y = M.reshape(-1, 1), Y.reshape(-1, 1)
        return -(m - y) / lik_v, -0.5 * np.ones_like(m) / lik_v


class HetGaussian(object):
    """likelihoods/hetgaussian.py:46-73 (analytic, 2 latent functions)."""
    name = "HetGaussian"
    dims = (1, 2, 1)

    def get_metadata(self):
        return self.dims

    def _terms(self, Y, M, V, safe):
        m0, m1 = M[:, 0, None], M[:, 1, None]
        v0, v1 = V[:, 0, None], V[:, 1, None]
        precision = np.clip(safe_exp(-m1 + 0.5 * v1), -1e9, 1e9)
        sq = safe_square if safe else np.square
        squares = np.clip(sq(Y) + sq(m0) + v0 - 2 * m0 * Y, -1e9, 1e9)
        return m0, m1, precision, squares

    def var_exp(self, Y, M, V):
        m0, m1, precision, squares = self._terms(Y, M, V, True)
        return -0.5 * np.log(2 * np.pi) - 0.5 * m1 - 0.5 * precision * squares

    def var_exp_derivatives(self, Y, M, V):
        m0, m1, precision, squares = self._terms(Y, M, V, False)
        dm = np.hstack((precision * (Y - m0), 0.5 * (precision * squares - 1.0)))
        dv = np.hstack((-0.5 * precision, -0.25 * precision * squares))
        return dm, dv


class _OneD(object):
    """Shared 1-D GH-20 quadrature (bernoulli.py:82-111, poisson.py:66-95,
    exponential.py:70-99)."""
    dims = (1, 1, 1)
    T = 20

    def get_metadata(self):
        return self.dims

    def _f(self, M, V):
        gh_f, gh_w = gh_points(self.T)
        m, v = M.reshape(-1), V.reshape(-1)
        with np.errstate(invalid="ignore"):
            f = gh_f[None, :] * np.sqrt(2.0 * v[:, None]) + m[:, None]
        return f, gh_w / np.sqrt(np.pi)

    def var_exp(self, Y, M, V):
        f, w = self._f(M, V)
        y = np.tile(Y.reshape(-1, 1), (1, f.shape[1]))
        return self.logpdf(f, y).dot(w[:, None])

    def var_exp_derivatives(self, Y, M, V):
        f, w = self._f(M, V)
        y = np.tile(Y.reshape(-1, 1), (1, f.shape[1]))
        return self.dlogp_df(f, y).dot(w[:, None]), 0.5 * self.d2logp_df2(f, y).dot(w[:, None])


class Bernoulli(_OneD):
    """likelihoods/bernoulli.py:31-36,66-80."""
    name = "Bernoulli"

    def _p(self, f):
        ef = safe_exp(f)
        return ef, np.clip(ef / (1 + ef), 1e-9, 1.0 - 1e-9)

    def logpdf(self, f, y):
        _, p = self._p(f)
        return y * np.log(p) + (1 - y) * np.log(1 - p)

    def dlogp_df(self, f, y):
        ef, p = self._p(f)
        return ((y - p) / (1 - p)) * (1 / (1 + ef))

    def d2logp_df2(self, f, y):
        ef, p = self._p(f)
        return -p / (1 + ef)


class Poisson(_OneD):
    """likelihoods/poisson.py:31-34,56-64."""
    name = "Poisson"

    def logpdf(self, f, y):
        return -safe_exp(f) + y * f - gammaln(y + 1)

    def dlogp_df(self, f, y):
        return -safe_exp(f) + y

    def d2logp_df2(self, f, y):
        return -safe_exp(f)


class Exponential(_OneD):
    """likelihoods/exponential.py:28-32,58-68."""
    name = "Exponential"

    def _b(self, f):
        return np.clip(safe_exp(-f), 1e-9, 1e9)

    def logpdf(self, f, y):
        b = self._b(f)
        return -np.log(b) - y / b

    def dlogp_df(self, f, y):
        return 1 - y / self._b(f)

    def d2logp_df2(self, f, y):
        return -y / self._b(f)


class Categorical(object):
    """likelihoods/categorical.py:37-46,77-82,102-128,130-222 (K-1 functions,
    GH-10 tensor grid)."""
    name = "Categorical"
    T = 10

    def __init__(self, K):
        self.K = K
        self.dims = (1, K - 1, K - 1)   # categorical.py:287-291

    def get_metadata(self):
        return self.dims

    def onehot(self, y):
        Y1 = np.zeros((y.shape[0], self.K))
        for k in range(self.K):
            Y1[:, k, None] = (y.reshape(-1, 1) == k + 1).astype(int)
        return Y1

    def logpdf(self, F, y):
        """F (n, K-1), y (n,1) labels in 1..K -> (n,)."""
        Y1 = self.onehot(y)
        eF = safe_exp(F)
        den = 1 + eF.sum(1)[:, None]
        p = np.hstack((eF / den, 1 / den))
        p = np.clip(p, 1e-9, 1 - 1e-9)
        p = p / p.sum(1)[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            out = np.sum(np.where(Y1 > 0, Y1 * np.log(p), 0.0), axis=1)
        out[Y1.sum(1) != 1] = -np.inf  # label outside 1..K (scipy>=1.2: out of domain -> -inf)
        return out

    def dlogp_df(self, df, F, y):
        Y1 = self.onehot(y)
        return Y1[:, df, None] - Y1.sum(1)[:, None]  # quirk C-1: p renormalises to 1

    def d2logp_df2(self, df, F, y):
        Y1 = self.onehot(y)
        eF = safe_exp(F)
        den = 1 + eF.sum(1)[:, None]
        enum = safe_exp(F + F[:, df, None])
        enum[:, df] = safe_exp(F[:, df])
        p = enum.sum(1)[:, None] / safe_square(den)
        return -(Y1 * p).sum(1)[:, None]

    def _flatF(self, M, V):
        gh_f, gh_w = gh_points(self.T)
        N, D = M.shape
        grid = _grid(M, V, gh_f)
        shape = (N,) + (self.T,) * D
        F = np.stack([np.broadcast_to(g, shape).reshape(-1) for g in grid], axis=1)
        return F, gh_w, shape

    def var_exp(self, Y, M, V, chunk=2048):
        out = []
        D = M.shape[1]
        for s in range(0, M.shape[0], chunk):
            m, v, y = M[s:s + chunk], V[s:s + chunk], Y[s:s + chunk]
            with np.errstate(invalid="ignore"):
                F, gh_w, shape = self._flatF(m, v)
            yf = np.repeat(y.reshape(-1, 1), self.T ** D, axis=0)
            logp = self.logpdf(F, yf).reshape(shape)
            out.append(_contract(logp, gh_w, D, False))
        return np.concatenate(out)[:, None]

    def var_exp_derivatives(self, Y, M, V, chunk=2048):
        N, D = M.shape
        dm = np.empty((N, D))
        dv = np.empty((N, D))
        for s in range(0, N, chunk):
            m, v, y = M[s:s + chunk], V[s:s + chunk], Y[s:s + chunk]
            with np.errstate(invalid="ignore"):
                F, gh_w, shape = self._flatF(m, v)
            yf = np.repeat(y.reshape(-1, 1), self.T ** D, axis=0)
            for d in range(D):
                dm[s:s + chunk, d] = _contract(self.dlogp_df(d, F, yf).reshape(shape), gh_w, D, False)
                dv[s:s + chunk, d] = 0.5 * _contract(self.d2logp_df2(d, F, yf).reshape(shape), gh_w, D, False)
        return dm, dv


class _TwoD(object):
    """Shared GH-10x10 quadrature with the reference's double 1/sqrt(pi)
    normalisation (gamma.py:103-194, beta.py:106-197; quirk C-2)."""
    dims = (1, 2, 1)
    T = 10

    def get_metadata(self):
        return self.dims

    def _ab(self, M, V):
        gh_f, gh_w = gh_points(self.T)
        with np.errstate(invalid="ignore"):
            fa, fb = _grid(M, V, gh_f)
        a = np.clip(safe_exp(fa), 1e-9, 1e9)
        b = np.clip(safe_exp(fb), 1e-9, 1e9)
        return a, b, gh_w

    def var_exp(self, Y, M, V, chunk=8192):
        out = []
        for s in range(0, M.shape[0], chunk):
            a, b, gh_w = self._ab(M[s:s + chunk], V[s:s + chunk])
            y = Y[s:s + chunk].reshape(-1, 1, 1)
            out.append(_contract(self.logpdf_ab(a, b, y), gh_w, 2, True))
        return np.concatenate(out)[:, None]

    def var_exp_derivatives(self, Y, M, V, chunk=8192):
        dm, dv = [], []
        for s in range(0, M.shape[0], chunk):
            a, b, gh_w = self._ab(M[s:s + chunk], V[s:s + chunk])
            y = Y[s:s + chunk].reshape(-1, 1, 1)
            da, db = self.dlogp_ab(a, b, y)
            d2a, d2b = self.d2logp_ab(a, b, y)
            shp = np.broadcast_shapes(a.shape, b.shape)
            c = lambda A: _contract(np.broadcast_to(A, shp), gh_w, 2, True)
            dm.append(np.stack((c(da), c(db)), axis=1))
            dv.append(0.5 * np.stack((c(d2a), c(d2b)), axis=1))
        return np.concatenate(dm), np.concatenate(dv)

    # pointwise API in the reference's (F, y) form
    def _split(self, F):
        eF = safe_exp(F)
        return np.clip(eF[:, 0, None], 1e-9, 1e9), np.clip(eF[:, 1, None], 1e-9, 1e9)

    def logpdf(self, F, y):
        a, b = self._split(F)
        return self.logpdf_ab(a, b, y)

    def dlogp_df(self, F, y):
        a, b = self._split(F)
        return self.dlogp_ab(a, b, y)

    def d2logp_df2(self, F, y):
        a, b = self._split(F)
        return self.d2logp_ab(a, b, y)


class Gamma(_TwoD):
    """likelihoods/gamma.py:34-41,80-101."""
    name = "Gamma"

    def logpdf_ab(self, a, b, y):
        return -gammaln(a) + a * np.log(b) + (a - 1) * np.log(y) - b * y

    def dlogp_ab(self, a, b, y):
        return (-psi(a) + np.log(b) + np.log(y)) * a, a - b * y

    def d2logp_ab(self, a, b, y):
        return (-psi(a) - a * zeta(2, a) + np.log(b) + np.log(y)) * a, -y * b


class Beta(_TwoD):
    """likelihoods/beta.py:29-36,76-104."""
    name = "Beta"

    def logpdf_ab(self, a, b, y):
        return (a - 1) * np.log(y) + (b - 1) * np.log(1 - y) - betaln(a, b)

    def dlogp_ab(self, a, b, y):
        psi_ab = psi(a + b)
        return (psi_ab - psi(a) + np.log(y)) * a, (psi_ab - psi(b) + np.log(1 - y)) * b

    def d2logp_ab(self, a, b, y):
        psi_ab, zeta_ab = psi(a + b), zeta(2, a + b)
        return ((psi_ab + a * zeta_ab - psi(a) - a * zeta(2, a) + np.log(y)) * a,
                (psi_ab + b * zeta_ab - psi(b) - b * zeta(2, b) + np.log(1 - y)) * b)


_CLASSES = {c.name: c for c in (Gaussian, HetGaussian, Bernoulli, Poisson, Exponential, Categorical, Gamma, Beta)}


def make(spec):
    """spec = ('Gaussian', sigma) | ('Categorical', K) | ('Bernoulli',) | ..."""
    name = spec[0]
    if name == "Gaussian":
        return Gaussian(spec[1] if len(spec) > 1 else None)
    if name == "Categorical":
        return Categorical(spec[1])
    return _CLASSES[name]()


def generate_metadata(liks):
    """het_likelihood.py:24-44 -- integer index maps (bit-exact)."""
    y_index, f_index, d_index, p_index = [], [], [], []
    for t, lik in enumerate(liks):
        dim_y, dim_f, dim_p = lik.get_metadata()
        y_index += [t] * dim_y
        f_index += [t] * dim_f
        d_index += list(range(dim_f))
        p_index += [t] * dim_p
    return {"task_index": np.arange(len(liks)), "y_index": np.array(y_index, dtype=np.int_),
            "function_index": np.array(f_index, dtype=np.int_), "d_index": np.array(d_index, dtype=np.int_),
            "pred_index": np.array(p_index, dtype=np.int_)}
