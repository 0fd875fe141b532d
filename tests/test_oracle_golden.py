"""CPU: the travelling numpy restatement (oracle/) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/, see oracle/make_golden.py).  Tolerances are fp64 round-off."""
import numpy as np
import pytest

import golden_util as gu
from oracle import diag_oracle, likelihoods_np as lk, verbatim

TOL = 1e-9


def rel(a, b):
    b = np.asarray(b, dtype=float)
    return np.max(np.abs(np.asarray(a) - b)) / max(np.max(np.abs(b)), 1e-300)


@pytest.mark.parametrize("name", gu.CASES)
def test_diag_oracle_matches_reference(name):
    prob, g = gu.load_case(name)
    o = diag_oracle.elbo_and_grads(prob, want_rows=True)
    assert abs(o["log_marginal"][0, 0] - g["log_marginal"][0, 0]) <= 1e-11 * abs(g["log_marginal"][0, 0])
    assert rel(np.hstack(o["dL_dmu_u"]), g["dL_dmu_u"]) < TOL
    assert rel(np.hstack(o["dL_dL_u"]), g["dL_dL_u"]) < TOL
    assert rel(np.stack(o["dL_dKmm"]), g["dL_dKmm"]) < TOL
    assert rel(o["d_rbf"], g["d_rbf"]) < TOL
    assert rel(o["dW"], g["dW"]) < TOL
    assert rel(o["dkappa"], g["dkappa"]) < TOL
    assert rel(o["dZ"], g["dZ"]) < TOL
    meta = lk.generate_metadata([lk.make(s) for s in prob["lik_specs"]])
    for d in range(prob["J"]):
        t, f = int(meta["function_index"][d]), int(meta["d_index"][d])
        assert rel(o["rows"]["m"][t][:, f], g["m_fd_%d" % d][:, 0]) < TOL
        assert rel(o["rows"]["v"][t][:, f], g["v_fd_%d" % d][:, 0]) < 1e-8


@pytest.mark.parametrize("name", gu.CASES)
def test_metadata_bit_exact(name):
    prob, g = gu.load_case(name)
    meta = lk.generate_metadata([lk.make(s) for s in prob["lik_specs"]])
    for k in ("task_index", "y_index", "function_index", "d_index", "pred_index"):
        assert np.array_equal(np.asarray(meta[k]).ravel(), g["meta_" + k].ravel()), k


@pytest.mark.parametrize("tag", gu.LIK_TAGS)
def test_likelihoods_np_match_reference(tag):
    g = gu.load_likelihoods()
    lik = lk.make(gu.LIK_SPECS[tag])
    Y, M, V = g[tag + "_Y"], g[tag + "_M"], g[tag + "_V"]
    ve = lik.var_exp(Y, M, V)
    dm, dv = lik.var_exp_derivatives(Y, M, V)
    assert rel(ve, g[tag + "_ve"]) < 1e-11
    assert rel(dm, g[tag + "_dm"]) < 1e-11
    assert rel(dv, g[tag + "_dv"]) < 1e-11


def test_known_answers():
    # Gaussian var_exp is the closed form of E[log N(y | f, sigma^2)]
    rng = np.random.default_rng(0)
    y, m, v = rng.normal(size=(50, 1)), rng.normal(size=(50, 1)), rng.uniform(0.1, 2, (50, 1))
    s = 0.7
    ve = lk.Gaussian(s).var_exp(y, m, v)
    x, w = np.polynomial.hermite.hermgauss(40)
    f = x[None, :] * np.sqrt(2 * v) + m
    quad = ((-0.5 * np.log(2 * np.pi * s * s) - 0.5 * (y - f) ** 2 / (s * s)) * w[None, :] / np.sqrt(np.pi)).sum(1)
    assert np.allclose(ve[:, 0], quad, rtol=1e-12, atol=1e-12)
    # KL = 0 when q(u) = p(u): m = 0, L_u = chol(K_uu)
    from oracle import synth
    prob = synth.make_problem([("Gaussian", 0.5)], 30, 10, 2, seed=3)
    Kuu, Luu, _, _ = diag_oracle.latent_funs_cov(prob["Z"], prob["rbf_var"], prob["rbf_ls"], 2, 1)
    prob["m_u"] = np.zeros_like(prob["m_u"])
    prob["L_u"] = np.stack([diag_oracle.pack_lower(Luu[q]) for q in range(2)], axis=1)
    o = diag_oracle.elbo_and_grads(prob)
    assert abs(o["KL"]) < 1e-9
    # packed-lower <-> dense is the identity on packed vectors (bit-exact)
    flat = rng.normal(size=(55,))
    assert np.array_equal(diag_oracle.pack_lower(diag_oracle.unpack_lower(flat, 10)), flat)


def test_shard_sum_identity():
    """Sum of per-shard data terms == unsharded (the multi-GPU reduction identity)."""
    prob, g = gu.load_case("cfg3_small")
    full = diag_oracle.elbo_and_grads(prob)
    N = [x.shape[0] for x in prob["X"]]
    acc = None
    for r in range(3):
        sl = [slice(n * r // 3, n * (r + 1) // 3) for n in N]
        o = diag_oracle.elbo_and_grads(prob, row_slices=sl)
        part = np.concatenate([o["VE_sum"], o["dVE_dmu"].ravel(), o["dVE_dS"].ravel(), o["sdv"], o["sma"].ravel(), o["svc"].ravel()])
        acc = part if acc is None else acc + part
    ref = np.concatenate([full["VE_sum"], full["dVE_dmu"].ravel(), full["dVE_dS"].ravel(), full["sdv"], full["sma"].ravel(), full["svc"].ravel()])
    assert rel(acc, ref) < 1e-12


@pytest.mark.skipif(not verbatim.available(), reason="reference tree not present (GPU box)")
def test_golden_regenerates_from_reference():
    """The committed fixtures are what the unmodified reference produces here."""
    prob, g = gu.load_case("cfg3_small")
    lm, grads, ex = verbatim.run_inference(prob)
    assert abs(lm[0, 0] - g["log_marginal"][0, 0]) <= 1e-12 * abs(lm[0, 0])
    assert rel(np.hstack(grads["dL_dmu_u"]), g["dL_dmu_u"]) < 1e-12
