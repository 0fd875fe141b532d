"""CPU: closed-form known-answer tests pinning the GPy stand-in (oracle/gpy_standin.py) that lets the reference's
files run unmodified (GPy 1.9.5 is not installable here)."""
import numpy as np

from oracle import gpy_standin as gpy


def test_rbf_formula_and_gradients():
    rng = np.random.default_rng(1)
    X, X2 = rng.normal(size=(7, 2)), rng.normal(size=(5, 2))
    k = gpy.RBF(2, variance=1.7, lengthscale=0.6)
    K = np.asarray(k.K(X, X2))
    ref = 1.7 * np.exp(-0.5 * ((X[:, None, :] - X2[None, :, :]) ** 2).sum(-1) / 0.36)
    assert np.allclose(K, ref, rtol=1e-12)
    assert np.allclose(np.diag(np.asarray(k.K(X, X))), 1.7)
    G = rng.normal(size=K.shape)
    k.update_gradients_full(G, X, X2)
    g = np.array(k.gradient, dtype=float).ravel()
    eps = 1e-6
    kp = gpy.RBF(2, variance=1.7 + eps, lengthscale=0.6)
    lp = gpy.RBF(2, variance=1.7, lengthscale=0.6 + eps)
    fd_v = ((np.asarray(kp.K(X, X2)) - K) * G).sum() / eps
    fd_l = ((np.asarray(lp.K(X, X2)) - K) * G).sum() / eps
    assert np.allclose(g, [fd_v, fd_l], rtol=1e-4)
    gx = np.asarray(k.gradients_X(G, X, X2))
    Xp = X.copy()
    Xp[2, 1] += eps
    fd_x = ((np.asarray(k.K(Xp, X2)) - K) * G).sum() / eps
    assert np.allclose(gx[2, 1], fd_x, rtol=1e-4)


def test_linalg_identities():
    rng = np.random.default_rng(2)
    A = rng.normal(size=(9, 9))
    A = A.dot(A.T) + 9 * np.eye(9)
    L = gpy.jitchol(A)
    assert np.allclose(np.tril(L).dot(np.tril(L).T), A)
    Ai, _ = gpy.dpotri(np.asfortranarray(L))
    assert np.allclose(Ai.dot(A), np.eye(9), atol=1e-10)
    B = rng.normal(size=(9, 3))
    X, _ = gpy.dpotrs(L, B)
    assert np.allclose(A.dot(X), B)
    # jitter only on failure
    S = np.ones((4, 4))
    Lj = gpy.jitchol(S)
    assert np.all(np.isfinite(Lj))


def test_choleskies_roundtrip_bit_exact():
    rng = np.random.default_rng(3)
    flat = rng.normal(size=(15, 2))
    tri = gpy.flat_to_triang(flat)
    assert tri.shape == (2, 5, 5) and np.all(np.triu(tri[0], 1) == 0)
    ii, jj = np.tril_indices(5)
    assert np.array_equal(tri[1][ii, jj], flat[:, 1])
    assert np.array_equal(gpy.triang_to_flat(tri), flat)


def test_gh_points_and_safe_ops():
    lik = gpy.Likelihood(gpy.Identity(), "x")
    x, w = lik._gh_points(20)
    xr, wr = np.polynomial.hermite.hermgauss(20)
    assert np.array_equal(x, xr) and np.array_equal(w, wr)
    assert np.isfinite(gpy.safe_exp(1e4)) and gpy.safe_exp(1.0) == np.exp(1.0)
    assert np.isfinite(gpy.safe_square(1e200))
