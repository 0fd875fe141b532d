"""GPU bring-up of the tensor-core path: one case per process (a trapped kernel poisons the CUDA context).

  python tools/tc_check.py small <case>        engine(tc) vs CPU oracle, every block            (cases of stagecheck.py)
  python tools/tc_check.py scale <cfg> <N>     engine(tc) vs engine(fp64) on the GPU at N rows/task + phase timings
  python tools/tc_check.py time  <cfg> <N>     phase timings only (full / ve / elbo)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from oracle import synth  # noqa: E402
import parity_util as pu  # noqa: E402
from stagecheck import CASES  # noqa: E402

GRADS = ("dL_dmu_u", "dL_dL_u", "dL_dKmm", "d_rbf", "dW", "dkappa", "dZ")


def small(name, prec="tc"):
    c = dict(CASES[name])
    prob = synth.make_problem(c.pop("liks"), c.pop("N"), c.pop("M"), c.pop("Q"), Xdim=c.pop("Xdim"), seed=7, **c)
    t0 = time.time()
    err, out, o = pu.compare(prob, prec)
    st = err.pop("_status")
    print("CASE %s %s (%.1fs) elbo=%.10g oracle=%.10g status=%s" % (name, prec, time.time() - t0, out["log_marginal"][0, 0],
                                                                   o["log_marginal"][0, 0], st))
    print("   " + "  ".join("%s=%.2e" % (k, v) for k, v in err.items()))


def timed(eng, p, what, reps=3):
    eng.enable_timing(True)
    best = None
    for _ in range(reps):
        t0 = time.time()
        out = eng.evaluate(p, what=what, want_dKmm=(what == "full"))
        dt = time.time() - t0
        tm = eng.last_timing()
        tm["wall_ms"] = dt * 1e3
        if best is None or tm["wall_ms"] < best["wall_ms"]:
            best = tm
    return out, best


def make_cfg(cfg, N):
    """cfg1..cfg4 of BASELINE.json, or 'sweepM<M>' = cfg5's inducing-point sweep (cfg2's likelihood list)."""
    if cfg.startswith("sweepM"):
        return synth.make_problem([("Gaussian", 0.5), ("Bernoulli",), ("Poisson",)], N, int(cfg[6:]), 3, 1, seed=1239)
    return synth.make_config(cfg, N=N)


def scale(cfg, N, ref_prec="fp64", whats=("full",)):
    prob = make_cfg(cfg, N)
    p = pu.params_of(prob)
    eng = pu.make_engine(prob, "tc")
    outs = {}
    for what in whats:
        outs[what], tm = timed(eng, p, what)
        print("TIME %s N=%d tc %-5s %s" % (cfg, N, what, "  ".join("%s=%.2f" % (k, v) for k, v in tm.items())))
        sys.stdout.flush()
    eng.close()
    if ref_prec is None:
        return
    ref = pu.make_engine(prob, ref_prec)
    o, tm = timed(ref, p, "full", reps=1)
    print("TIME %s N=%d %s full %s" % (cfg, N, ref_prec, "  ".join("%s=%.2f" % (k, v) for k, v in tm.items())))
    ref.close()
    out = outs["full"]
    e = abs(out["log_marginal"][0, 0] - o["log_marginal"][0, 0]) / abs(o["log_marginal"][0, 0])
    print("PARITY %s N=%d tc vs %s: elbo=%.12g ref=%.12g rel=%.2e  " % (cfg, N, ref_prec, out["log_marginal"][0, 0],
                                                                      o["log_marginal"][0, 0], e) +
          "  ".join("%s=%.2e" % (k, pu.relerr(out[k], o[k])) for k in GRADS))
    if "ve" in outs:
        print("PARITY %s N=%d tc ve vs %s: " % (cfg, N, ref_prec) + "  ".join("%s=%.2e" % (k, pu.relerr(outs["ve"][k], o[k])) for k in ("dL_dmu_u", "dL_dL_u")))
    print("D_RBF tc ", np.array2string(out["d_rbf"].ravel(), precision=6), " ref", np.array2string(o["d_rbf"].ravel(), precision=6))
    if os.environ.get("TC_CHECK_FP32"):
        e32 = pu.make_engine(prob, "fp32")
        o32, tm = timed(e32, p, "full", reps=1)
        e32.close()
        e = abs(o32["log_marginal"][0, 0] - o["log_marginal"][0, 0]) / abs(o["log_marginal"][0, 0])
        print("PARITY %s N=%d fp32 vs %s: rel=%.2e  " % (cfg, N, ref_prec, e) +
              "  ".join("%s=%.2e" % (k, pu.relerr(o32[k], o[k])) for k in GRADS))


def determinism(cfg, N):
    """Bit-reproducibility of repeated evaluations and of the ELBO across what-levels."""
    prob = synth.make_config(cfg, N=N)
    p = pu.params_of(prob)
    eng = pu.make_engine(prob, "tc")
    vals = {}
    for what in ("full", "full", "ve", "ve", "elbo", "elbo", "full"):
        o = eng.evaluate(p, what=what)
        vals.setdefault(what, []).append((repr(float(o["log_marginal"][0, 0])), [repr(float(v)) for v in o["VE"]],
                                          None if what == "elbo" else float(np.abs(o["dL_dL_u"]).sum())))
    for k, v in vals.items():
        for x in v:
            print("DET", k, x)
    eng.close()


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "det":
        determinism(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    if mode == "small":
        small(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "tc")
    elif mode == "scale":
        scale(sys.argv[2], int(sys.argv[3]), whats=("full", "ve", "elbo"))
    elif mode == "time":
        scale(sys.argv[2], int(sys.argv[3]), ref_prec=None, whats=("full", "ve", "elbo"))
