"""Shared helpers of the parity tests: run the CUDA engine and the CPU oracle on the same problem dict."""
import numpy as np


def params_of(problem):
    keys = ("Z", "m_u", "L_u", "rbf_var", "rbf_ls", "W", "kappa")
    p = {k: np.ascontiguousarray(problem[k], dtype=np.float64) for k in keys}
    p["batch_scale"] = np.asarray(problem.get("batch_scale") or [1.0] * len(problem["Y"]), dtype=np.float64)
    for k in ("W_chain", "kappa_chain"):
        if problem.get(k) is not None:
            p[k] = np.ascontiguousarray(problem[k], dtype=np.float64)
    return p


def make_engine(problem, precision, **kw):
    from hetmogp_b200 import Engine
    eng = Engine(problem["lik_specs"], problem["M"], problem["Q"], problem["Xdim"], precision=precision, **kw)
    eng.set_data(problem["X"], problem["Y"])
    return eng


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def compare(problem, precision, oracle_out=None, rows=True):
    """Return ({name: rel-inf-norm error vs oracle}, engine outputs, oracle outputs)."""
    from oracle import diag_oracle
    if oracle_out is None:
        oracle_out = diag_oracle.elbo_and_grads(problem, want_rows=rows, W_chain=problem.get("W_chain"),
                                                kappa_chain=problem.get("kappa_chain"))
    o = oracle_out
    eng = make_engine(problem, precision)
    out = eng.evaluate(params_of(problem), what="full", want_dKmm=True)
    Q = problem["Q"]
    err = {}
    err["elbo"] = abs(out["log_marginal"][0, 0] - o["log_marginal"][0, 0]) / abs(o["log_marginal"][0, 0])
    err["VE"] = relerr(out["VE"], o["VE_sum"])
    err["KL"] = abs(out["KL"][0] - o["KL"]) / max(abs(o["KL"]), 1e-300)
    err["dL_dmu_u"] = relerr(out["dL_dmu_u"], np.hstack(o["dL_dmu_u"]))
    err["dL_dL_u"] = relerr(out["dL_dL_u"], np.hstack(o["dL_dL_u"]))
    err["dL_dKmm"] = relerr(out["dL_dKmm"], np.stack(o["dL_dKmm"]))
    err["d_rbf"] = relerr(out["d_rbf"], o["d_rbf"])
    err["dW"] = relerr(out["dW"], o["dW"])
    err["dkappa"] = relerr(out["dkappa"], o["dkappa"])
    err["dZ"] = relerr(out["dZ"], o["dZ"])
    Kuu, Luu, Kuui = eng.kuu()
    err["Kuu"] = relerr(Kuu, o["Kuu"])
    err["Luu"] = relerr(Luu, o["Luu"])
    err["Kuui"] = relerr(Kuui, o["Kuui"])
    if rows and "rows" in o:
        for t in range(len(problem["Y"])):
            r = eng.rows(t)
            for k in ("m", "v", "ve", "dm", "dv"):
                err["row_%s[%d]" % (k, t)] = relerr(r[k], o["rows"][k][t])
    err["_status"] = eng.status
    eng.close()
    return err, out, o
