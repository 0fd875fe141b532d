"""Seeded synthetic problems for the HetMOGP hot path.

BENCH / TEST INPUT GENERATOR.  Pure numpy; no arithmetic of the hot path, no reference or oracle access.

A *problem* is a dict of plain numpy arrays in the layout of the reference's
inference() arguments (svmogp_inf.py:23-24):

  X      list[T] of (N_t, Xdim) f64        Y   list[T] of (N_t, 1) f64
  Z      (M, Q*Xdim) f64 -- column block q holds latent q's inducing inputs
         (util.py:197; svmogp.py:52 tiles one Z)
  m_u    (M, Q) f64                        L_u (M(M+1)/2, Q) packed lower, row-major
  rbf_var, rbf_ls (Q,)                     W, kappa (J, Q)   [W_q = W[:, q:q+1]]
  lik_specs list[T] of tuples              batch_scale list[T] of floats
"""
import numpy as np

def _dim_f(spec):
    """Latent functions per likelihood (get_metadata of likelihoods/*.py)."""
    name = spec[0]
    if name in ("HetGaussian", "Gamma", "Beta"):
        return 2
    if name == "Categorical":
        return int(spec[1]) - 1
    return 1


CONFIGS = {
    # BASELINE.json configs[0..3]; cfg5 is a sweep over M with cfg2's list
    "cfg1": dict(N=200, M=20, Q=2, Xdim=1, liks=[("HetGaussian",), ("Bernoulli",), ("Categorical", 3)]),
    "cfg2": dict(N=100000, M=200, Q=3, Xdim=1, liks=[("Gaussian", 0.5), ("Bernoulli",), ("Poisson",)]),
    "cfg3": dict(N=1000000, M=500, Q=3, Xdim=1,
                 liks=[("HetGaussian",), ("Bernoulli",), ("Categorical", 4), ("Gamma",), ("Beta",)]),
    "cfg4": dict(N=500000, M=1000, Q=2, Xdim=2, liks=[("Categorical", 4), ("Gaussian", 0.5)]),
}


def inducing_grid(M, Xdim, rng):
    """1-D: linspace(0,1,M) (demo.ipynb:227).  2-D: ceil(sqrt(M))^2 jittered grid
    truncated to M.  Returns (Z (M,Xdim), grid spacing h)."""
    if Xdim == 1:
        return np.linspace(0.0, 1.0, M)[:, None], 1.0 / (M - 1)
    g = int(np.ceil(M ** (1.0 / Xdim)))
    axes = [np.linspace(0.0, 1.0, g)] * Xdim
    Z = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, Xdim)
    h = 1.0 / (g - 1)
    Z = Z + 0.05 * h * rng.standard_normal(Z.shape)
    return Z[:M].copy(), h


def _latent_draw(rng, Q):
    """Smooth bounded stand-ins for u_q(x) in the style of util.py:21-34
    (three random sinusoids per latent)."""
    return dict(amp=rng.uniform(0.3, 0.8, (Q, 3)), freq=rng.uniform(1.0, 3.0, (Q, 3)),
                shift=rng.uniform(0.0, 2.0, (Q, 3)))


def _latent_eval(lat, X):
    s = X.sum(axis=1)
    U = np.empty((X.shape[0], lat["amp"].shape[0]))
    for q in range(U.shape[1]):
        a, f, p = lat["amp"][q], lat["freq"][q], lat["shift"][q]
        U[:, q] = a[0] * np.cos(f[0] * np.pi * s + p[0] * np.pi) - a[1] * np.sin(2 * f[1] * np.pi * s + p[1] * np.pi) \
            + a[2] * np.cos(4 * f[2] * np.pi * s + p[2] * np.pi)
    return U


def sample_outputs(spec, F, rng):
    """Draw Y ~ p(y | f) for one task; Y is stored as float (N,1)."""
    name = spec[0]
    if name == "Gaussian":
        sigma = spec[1] if len(spec) > 1 and spec[1] is not None else 0.5
        y = F[:, 0] + sigma * rng.standard_normal(F.shape[0])
    elif name == "HetGaussian":
        y = F[:, 0] + np.exp(0.5 * F[:, 1]) * rng.standard_normal(F.shape[0])
    elif name == "Bernoulli":
        y = (rng.uniform(size=F.shape[0]) < 1.0 / (1.0 + np.exp(-F[:, 0]))).astype(float)
    elif name == "Poisson":
        y = rng.poisson(np.exp(F[:, 0])).astype(float)
    elif name == "Exponential":
        y = rng.exponential(np.exp(-F[:, 0]))
        y = np.maximum(y, 1e-12)
    elif name == "Categorical":
        eF = np.exp(F)
        den = 1.0 + eF.sum(1, keepdims=True)
        p = np.hstack((eF / den, 1.0 / den))
        c = np.cumsum(p, axis=1)
        u = rng.uniform(size=(F.shape[0], 1))
        y = (1 + (u > c[:, :-1]).sum(1)).astype(float)  # labels 1..K (categorical.py:77-87)
    elif name == "Gamma":
        y = rng.gamma(np.exp(F[:, 0]), 1.0 / np.exp(F[:, 1]))
        y = np.maximum(y, 1e-12)
    elif name == "Beta":
        y = np.clip(rng.beta(np.exp(F[:, 0]), np.exp(F[:, 1])), 1e-6, 1.0 - 1e-6)
    else:
        raise ValueError(name)
    return y[:, None]


def make_problem(liks, N, M, Q, Xdim=1, seed=1234, ls_factor=(1.0, 1.15, 1.3), batch_scale=None,
                 kappa_scale=0.0):
    """Build a seeded problem: X_t ~ U[0,1]^Xdim drawn per
    task, Z a regular grid, RBF variances (1.0,0.7,1.3), lengthscales
    ls_factor*h so cond(K_uu) stays ~1e2..1e4, W ~ 0.3*random_W_kappas law
    (util.py:96-98), kappa=kappa_scale*U (reference: zeros, util.py:100-102),
    m_u ~ 0.1 N(0,1), L_u = 0.5 I + tril(0.3 N(0,1)/sqrt(M))."""
    rng = np.random.default_rng(seed)
    T = len(liks)
    Ns = [N] * T if np.isscalar(N) else list(N)
    meta = {"function_index": np.concatenate([np.full(_dim_f(s), t, dtype=np.int64) for t, s in enumerate(liks)])}
    J = meta["function_index"].shape[0]
    Zg, h = inducing_grid(M, Xdim, rng)
    Z = np.tile(Zg, (1, Q))
    rbf_var = np.array((1.0, 0.7, 1.3, 0.9, 1.1, 0.8)[:Q])
    rbf_ls = np.array((tuple(ls_factor) * 3)[:Q]) * h
    p = rng.binomial(1, 0.5, (J, Q))
    W = 0.3 * (p * rng.normal(0.5, 0.5, (J, Q)) - (p - 1) * rng.normal(-0.5, 0.5, (J, Q)))
    kappa = kappa_scale * rng.uniform(0.1, 0.5, (J, Q))
    m_u = 0.1 * rng.standard_normal((M, Q))
    ii, jj = np.tril_indices(M)
    L_u = np.empty((M * (M + 1) // 2, Q))
    for q in range(Q):
        L = 0.5 * np.eye(M) + np.tril(0.3 * rng.standard_normal((M, M)) / np.sqrt(M))
        L_u[:, q] = L[ii, jj]
    lat = _latent_draw(rng, Q)
    X, Y = [], []
    for t in range(T):
        Xt = rng.uniform(0.0, 1.0, (Ns[t], Xdim))
        U = _latent_eval(lat, Xt)
        ds = np.nonzero(meta["function_index"] == t)[0]
        F = U.dot(W[ds].T) * 3.0  # modest dynamic range; keeps exp-link likelihoods unsaturated
        X.append(Xt)
        Y.append(sample_outputs(liks[t], F, rng))
    return dict(X=X, Y=Y, Z=Z, m_u=m_u, L_u=L_u, rbf_var=rbf_var, rbf_ls=rbf_ls, W=W, kappa=kappa,
                lik_specs=list(liks), Q=Q, M=M, T=T, J=J, Xdim=Xdim, seed=seed,
                batch_scale=[1.0] * T if batch_scale is None else list(batch_scale))


def make_config(name, N=None, seed=None):
    """Problem for a BASELINE.json config; N overrides rows per task (bounded
    samples of the same workload)."""
    c = CONFIGS[name]
    idx = int(name[3:])
    return make_problem(c["liks"], c["N"] if N is None else N, c["M"], c["Q"], c["Xdim"],
                        seed=(1234 + idx) if seed is None else seed)


def subsample(problem, n_rows):
    """First n_rows of every task (rows are i.i.d., so a prefix is a sample)."""
    p = dict(problem)
    p["X"] = [x[:n_rows] for x in problem["X"]]
    p["Y"] = [y[:n_rows] for y in problem["Y"]]
    return p
