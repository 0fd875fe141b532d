"""Write tests/golden/*.npz from the UNMODIFIED reference (container only; needs /root/reference).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Run:  python -m oracle.make_golden
The fixtures pin (a) the travelling numpy restatement (oracle/diag_oracle.py, likelihoods_np.py) and (b) the CUDA
engine, against outputs of the reference's own code: SVMOGPInf.inference (hetmogp/svmogp_inf.py:23-109), the
likelihood classes (likelihoods/*.py) and HetLikelihood.generate_metadata (hetmogp/het_likelihood.py:24-44);
parameters_changed (hetmogp/svmogp.py:85-166) is applied through its line-by-line restatement
oracle/params_changed.py on the reference's dense gradients dict.
"""
import os
import warnings

import numpy as np

from . import params_changed, synth, verbatim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

ALL = [("HetGaussian",), ("Bernoulli",), ("Categorical", 3), ("Gamma",), ("Beta",), ("Poisson",), ("Gaussian", 0.5),
       ("Exponential",), ("Categorical", 4)]

INFERENCE_CASES = {
    # name: make_problem kwargs
    "cfg1_toy": dict(liks=[("HetGaussian",), ("Bernoulli",), ("Categorical", 3)], N=200, M=20, Q=2, Xdim=1, seed=1235),
    "all_liks": dict(liks=ALL, N=[50, 60, 40, 30, 45, 50, 20, 33, 25], M=12, Q=3, Xdim=1, seed=11,
                     batch_scale=[1.0, 2.0, 1.5, 1.0, 1.0, 3.0, 1.0, 1.0, 1.25]),
    "cfg2_small": dict(liks=[("Gaussian", 0.5), ("Bernoulli",), ("Poisson",)], N=120, M=24, Q=3, Xdim=1, seed=1236),
    "cfg3_small": dict(liks=[("HetGaussian",), ("Bernoulli",), ("Categorical", 4), ("Gamma",), ("Beta",)], N=60, M=16,
                       Q=3, Xdim=1, seed=1237),
    "cfg4_small": dict(liks=[("Categorical", 4), ("Gaussian", 0.5)], N=[90, 70], M=16, Q=2, Xdim=2, seed=1238,
                       kappa_scale=1.0),
}


def problem_from_case(c):
    c = dict(c)
    return synth.make_problem(c.pop("liks"), c.pop("N"), c.pop("M"), c.pop("Q"), Xdim=c.pop("Xdim"), **c)


def inference_golden(name, case):
    prob = problem_from_case(case)
    lm, grads, ex = verbatim.run_inference(prob)
    pc = params_changed.assemble(grads, prob, ex["Y_metadata"])
    Q, J = prob["Q"], prob["J"]
    out = dict(log_marginal=np.asarray(lm),
               dL_dmu_u=np.hstack(grads["dL_dmu_u"]), dL_dL_u=np.hstack(grads["dL_dL_u"]),
               dL_dKmm=np.stack(grads["dL_dKmm"]),
               d_rbf=pc["rbf"], dW=pc["W"], dkappa=pc["kappa"], dZ=pc["Z"])
    for d in range(J):
        out["m_fd_%d" % d] = ex["m_fd"][d]
        out["v_fd_%d" % d] = ex["v_fd"][d]
        for q in range(Q):
            out["dL_dKmn_%d_%d" % (q, d)] = np.asarray(grads["dL_dKmn"][q][d])
            out["dL_dKdiag_%d_%d" % (q, d)] = np.asarray(grads["dL_dKdiag"][q][d])
    for k, v in ex["Y_metadata"].items():
        out["meta_" + k] = np.asarray(v)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "inference_%s.npz" % name), **out)
    return out


def likelihood_golden():
    """var_exp / var_exp_derivatives / logpdf / dlogp_df / d2logp_df2 of every reference likelihood class."""
    ns = verbatim.load()
    rng = np.random.default_rng(2024)
    out = {}
    n = 64
    for spec in ALL:
        lik = verbatim.make_likelihood(ns, spec)
        tag = spec[0] + (str(spec[1]) if spec[0] == "Categorical" else "")
        _, F, _ = lik.get_metadata()
        M = rng.normal(0.0, 1.0, (n, F))
        V = rng.uniform(0.05, 1.5, (n, F))
        if spec[0] == "Gaussian" or spec[0] == "HetGaussian":
            Y = rng.normal(0, 1, (n, 1))
        elif spec[0] == "Bernoulli":
            Y = rng.integers(0, 2, (n, 1)).astype(float)
        elif spec[0] == "Poisson":
            Y = rng.poisson(2.0, (n, 1)).astype(float)
        elif spec[0] == "Categorical":
            Y = rng.integers(1, spec[1] + 1, (n, 1)).astype(float)
        elif spec[0] == "Beta":
            Y = rng.uniform(0.02, 0.98, (n, 1))
        else:
            Y = rng.gamma(2.0, 1.0, (n, 1)) + 1e-3
        # include a few extreme rows (clip regions)
        M[:4] *= 8.0
        V[:4] *= 4.0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ve = lik.var_exp(Y, M, V)
            dm, dv = lik.var_exp_derivatives(Y, M, V)
        out[tag + "_Y"], out[tag + "_M"], out[tag + "_V"] = Y, M, V
        out[tag + "_ve"], out[tag + "_dm"], out[tag + "_dv"] = np.asarray(ve).reshape(n, 1), np.asarray(dm).reshape(n, F), np.asarray(dv).reshape(n, F)
        # prediction (svmogp.py:340-378): predictive moments on the instance that has already run var_exp (so the cached
        # Gauss-Hermite table is the training one) and the seeded Monte-Carlo log predictive
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if spec[0] == "Gaussian":
                pm, pv = lik.predictive(M, V, None)
            else:
                pm, pv = lik.predictive(M, V)
            out[tag + "_pm"], out[tag + "_pv"] = np.asarray(pm).reshape(n, -1), np.asarray(pv).reshape(n, -1)
            if hasattr(lik, "log_predictive"):
                np.random.seed(7)
                out[tag + "_lp"] = np.asarray(lik.log_predictive(Y, M, V, 40)).reshape(1)
        # pointwise at F = M
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if spec[0] == "Categorical":
                out[tag + "_logpdf"] = np.asarray(lik.logpdf(M, Y)).reshape(n)
                out[tag + "_dlogp"] = np.hstack([np.asarray(lik.dlogp_df(d, M, Y)).reshape(n, 1) for d in range(F)])
                out[tag + "_d2logp"] = np.hstack([np.asarray(lik.d2logp_df2(d, M, Y)).reshape(n, 1) for d in range(F)])
            elif spec[0] in ("Gamma", "Beta"):
                out[tag + "_logpdf"] = np.asarray(lik.logpdf(M, Y)).reshape(n)
                out[tag + "_dlogp"] = np.hstack([np.asarray(a).reshape(n, 1) for a in lik.dlogp_df(M, Y)])
                out[tag + "_d2logp"] = np.hstack([np.asarray(a).reshape(n, 1) for a in lik.d2logp_df2(M, Y)])
            elif spec[0] in ("Bernoulli", "Poisson", "Exponential"):
                out[tag + "_logpdf"] = np.asarray(lik.logpdf(M, Y)).reshape(n)
                out[tag + "_dlogp"] = np.asarray(lik.dlogp_df(M, Y)).reshape(n, 1)
                out[tag + "_d2logp"] = np.asarray(lik.d2logp_df2(M, Y)).reshape(n, 1)
            elif spec[0] == "HetGaussian":
                out[tag + "_logpdf"] = np.asarray(lik.logpdf(M, Y[:, 0])).reshape(n)
            else:
                out[tag + "_logpdf"] = np.asarray(lik.logpdf(M, Y)).reshape(n)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "likelihoods.npz"), **out)
    return out


STREAM_CASES = [(200, 50), (203, 50), (7, 3), (5, 8), (1000, 64), (64, 64)]      # (n_samples, batch_size)


def util_golden():
    """Host helpers of hetmogp/util.py that produce integer / seeded outputs: the minibatch slice stream (util.py:52-72,
    two epochs), get_batch_scales (util.py:15-19), random_W_kappas (util.py:92-104) and the toy generators
    (util.py:21-50, 202-206) under fixed numpy seeds."""
    ns = verbatim.load()
    u = ns.util
    out = {}
    for n, bs in STREAM_CASES:
        sl = u.mini_slices(n, bs)
        out["mini_%d_%d" % (n, bs)] = np.array([[s.start, s.stop] for s in sl], dtype=np.int64)
        gen = u.draw_mini_slices(n, bs)
        seq = [next(gen) for _ in range(2 * len(sl) + 1)]
        out["draw_%d_%d" % (n, bs)] = np.array([[s.start, s.stop] for s in seq], dtype=np.int64)
        X = np.zeros((n, 1))
        out["len_%d_%d" % (n, bs)] = np.array([X[s].shape[0] for s in seq], dtype=np.int64)
        out["scale_%d_%d" % (n, bs)] = np.array([u.get_batch_scales([X], [X[s]])[0] for s in seq if X[s].shape[0] > 0])
    np.random.seed(101)
    W_list, kappa_list = u.random_W_kappas(3, 5, rank=1)
    out["rwk_W"] = np.hstack(W_list)
    out["rwk_kappa"] = np.stack(kappa_list, axis=1)
    np.random.seed(102)
    Xl = [np.linspace(0, 1, 17)[:, None], np.linspace(-1, 2, 9)[:, None]]
    tu = u.true_u_functions(Xl, 3)
    out["true_u_0"], out["true_u_1"] = tu[0], tu[1]
    liks = [verbatim.make_likelihood(ns, s) for s in (("HetGaussian",), ("Bernoulli",))]
    meta = ns.HetLikelihood(liks).generate_metadata()
    np.random.seed(103)
    W_list, _ = u.random_W_kappas(3, 3, rank=1)
    tf = u.true_f_functions(tu, W_list, 3, liks, meta)
    out["true_f_W"] = np.hstack(W_list)
    out["true_f_0"], out["true_f_1"] = tf[0], tf[1]
    np.random.seed(104)
    out["toy_U"] = u.generate_toy_U(np.linspace(0, 1, 11)[:, None], 4)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "util_streams.npz"), **out)
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    o = util_golden()
    print("util_streams: %d arrays" % len(o))
    for name, case in INFERENCE_CASES.items():
        o = inference_golden(name, case)
        print("inference_%s: log_marginal = %.12g" % (name, o["log_marginal"][0, 0]))
    o = likelihood_golden()
    print("likelihoods: %d arrays" % len(o))


if __name__ == "__main__":
    main()
