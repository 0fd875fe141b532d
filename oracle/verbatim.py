"""Run the reference's own hot-path files UNMODIFIED (container only).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  ``/root/reference`` does not
exist on the GPU box, so this module is used only by ``oracle/make_golden.py``
and by CPU tests that skip when the reference tree is absent.  Nothing is
copied: the files are imported from where they lie.

Imported verbatim: hetmogp/svmogp_inf.py, hetmogp/util.py,
hetmogp/het_likelihood.py, likelihoods/*.py.  hetmogp/svmogp.py cannot be
imported (it subclasses GPy.core.SparseGP, svmogp.py:16); its
parameters_changed (svmogp.py:85-166) is restated in oracle/params_changed.py.
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("HETMOGP_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hetmogp", "svmogp_inf.py"))


_cache = {}


def load():
    """Return a namespace with the reference modules (svmogp_inf, util,
    het_likelihood and the likelihood classes)."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from . import gpy_standin

    gpy_standin.install()
    # the reference's packages are called 'hetmogp' and 'likelihoods'
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib

        svmogp_inf = importlib.import_module("hetmogp.svmogp_inf")
        util = importlib.import_module("hetmogp.util")
        het = importlib.import_module("hetmogp.het_likelihood")
        liks = {}
        for mod, cls in (("gaussian", "Gaussian"), ("hetgaussian", "HetGaussian"), ("bernoulli", "Bernoulli"),
                         ("poisson", "Poisson"), ("categorical", "Categorical"), ("gamma", "Gamma"),
                         ("beta", "Beta"), ("exponential", "Exponential")):
            liks[cls] = getattr(importlib.import_module("likelihoods." + mod), cls)

    class NS(object):
        pass

    ns = NS()
    ns.svmogp_inf = svmogp_inf
    ns.util = util
    ns.het_likelihood = het
    ns.SVMOGPInf = svmogp_inf.SVMOGPInf
    ns.HetLikelihood = het.HetLikelihood
    ns.likelihoods = liks
    ns.gpy = gpy_standin
    _cache["ns"] = ns
    return ns


def make_likelihood(ns, spec):
    """spec = ('Gaussian', sigma) | ('Categorical', K) | ('Bernoulli',) ..."""
    name = spec[0]
    cls = ns.likelihoods[name]
    if name == "Gaussian":
        return cls(sigma=spec[1] if len(spec) > 1 else None)
    if name == "Categorical":
        return cls(K=spec[1])
    return cls()


def run_inference(problem):
    """Execute SVMOGPInf.inference (svmogp_inf.py:23) verbatim on a problem dict
    (see oracle/synth.py) and return (log_marginal, gradients, extras)."""
    import numpy as np

    ns = load()
    gpy = ns.gpy
    Q = problem["Q"]
    Xdim = problem["Xdim"]
    lik_list = [make_likelihood(ns, s) for s in problem["lik_specs"]]
    likelihood = ns.HetLikelihood(lik_list)
    Y_metadata = likelihood.generate_metadata()
    kern_list = [gpy.RBF(Xdim, variance=problem["rbf_var"][q], lengthscale=problem["rbf_ls"][q]) for q in range(Q)]
    B_list = [gpy.Coregionalize(Xdim, problem["J"], rank=1, W=problem["W"][:, q:q + 1], kappa=problem["kappa"][:, q])
              for q in range(Q)]
    inf = ns.SVMOGPInf()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        log_marginal, gradients, posteriors, posteriors_F = inf.inference(
            q_u_means=problem["m_u"].copy(), q_u_chols=problem["L_u"].copy(), X=problem["X"], Y=problem["Y"],
            Z=problem["Z"].copy(), kern_list=kern_list, likelihood=likelihood, B_list=B_list,
            Y_metadata=Y_metadata, batch_scale=problem.get("batch_scale"))
    extras = {"Y_metadata": Y_metadata, "kern_list": kern_list, "B_list": B_list, "likelihood": likelihood,
              "m_fd": [np.asarray(p.mean) for p in posteriors_F],
              "v_fd": [np.diag(np.asarray(p.covariance)).copy()[:, None] for p in posteriors_F]}
    return log_marginal, gradients, extras
