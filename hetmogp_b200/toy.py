"""Toy-data generators with the reference's names and random streams (hetmogp/util.py:21-50, 202-206).

Host-side input generation only (no arithmetic of the hot path).  Each function draws from ``numpy.random`` in the same
order and with the same shapes as the reference, so a seeded notebook cell produces the same data; the evaluation is
vectorised over the latent index instead of looping.
"""
import numpy as np


def true_u_functions(X_list, Q):
    """Q smooth latent functions per task: three random sinusoids each (util.py:21-34).  Draw order: amplitude (Q,3),
    frequency (Q,3), phase shift (Q,3), shared by all tasks."""
    amp = 0.5 + np.random.rand(Q, 3)                  # (1.5 - 0.5) * U + 0.5
    freq = 1.0 + 2.0 * np.random.rand(Q, 3)           # (3 - 1) * U + 1
    shift = 2.0 * np.random.rand(Q, 3)
    out = []
    for X in X_list:
        x = np.asarray(X, dtype=np.float64).reshape(X.shape[0], -1)[:, :1]          # (N, 1) against (Q,) -> (N, Q)
        u = (3.0 * amp[:, 0]) * np.cos(freq[:, 0] * np.pi * x + shift[:, 0] * np.pi) \
            - (2.0 * amp[:, 1]) * np.sin(2.0 * freq[:, 1] * np.pi * x + shift[:, 1] * np.pi) \
            + amp[:, 2] * np.cos(4.0 * freq[:, 2] * np.pi * x + shift[:, 2] * np.pi)
        out.append(u)
    return out


def true_f_functions(true_u, W_list, D, likelihood_list, Y_metadata):
    """Linear mix of the latent functions into every task's parameter functions (util.py:36-50):
    F_t[:, d_index[d]] = sum_q W_q[d] u_q for the output functions d of task t."""
    f_index = np.asarray(Y_metadata['function_index']).ravel()
    d_index = np.asarray(Y_metadata['d_index']).ravel()
    W = np.hstack([np.asarray(w, dtype=np.float64).reshape(-1, 1) for w in W_list])   # (D, Q)
    out = []
    for t, u in enumerate(true_u):
        _, n_f, _ = likelihood_list[t].get_metadata()
        F = np.zeros((u.shape[0], n_f))
        for q in range(W.shape[1]):                      # same accumulation order as the reference (q outer, d inner)
            for d in np.nonzero(f_index[:D] == t)[0]:
                F[:, d_index[d]] += W[d, q] * u[:, q]
        out.append(F)
    return out


def generate_toy_U(X, Q):
    """util.py:202-206: one random frequency/amplitude per latent, two random phases; draw order rand(1,Q), randn(1),
    randn(1)."""
    X = np.asarray(X, dtype=np.float64)
    arg = np.tile(X, (1, Q))
    rnd = np.tile(np.random.rand(1, Q), X.shape)
    p1 = np.random.randn(1)
    p2 = np.random.randn(1)
    return 2 * rnd * np.sin(10 * rnd * arg + p1) + 2 * rnd * np.cos(20 * rnd * arg + p2)
