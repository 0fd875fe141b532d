"""GPU bring-up: stage-wise comparison of the CUDA engine with the CPU oracle on several seeded problems."""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from oracle import synth  # noqa: E402
import parity_util as pu  # noqa: E402

ALL = [("HetGaussian",), ("Bernoulli",), ("Categorical", 3), ("Gamma",), ("Beta",), ("Poisson",), ("Gaussian", 0.5),
       ("Exponential",), ("Categorical", 4)]
CASES = {
    "toy": dict(liks=[("HetGaussian",), ("Bernoulli",), ("Categorical", 3)], N=200, M=20, Q=2, Xdim=1),
    "all": dict(liks=ALL, N=[150, 260, 140, 130, 145, 150, 120, 77, 64], M=40, Q=3, Xdim=1, batch_scale=[1, 2, 1.5, 1, 1, 3, 1, 1, 1.25]),
    "m300": dict(liks=[("Gaussian", 0.5), ("Bernoulli",), ("Poisson",)], N=3000, M=300, Q=3, Xdim=1),
    "x2": dict(liks=[("Categorical", 4), ("Gaussian", 0.5)], N=[1500, 901], M=100, Q=2, Xdim=2, kappa_scale=1.0),
    "pair_m200_x2": dict(liks=[("Gamma",), ("Beta",), ("Gaussian", 0.5)], N=[1500, 700, 3], M=200, Q=2, Xdim=2),
    "pair_m200_x2_old": dict(liks=[("Gamma",), ("Beta",), ("Gaussian", 0.5)], N=[1500, 700, 3], M=200, Q=2, Xdim=2),
    "pair_m700_x3": dict(liks=[("Bernoulli",), ("Poisson",)], N=[900, 130], M=700, Q=1, Xdim=3),
}
if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for name in names:
        c = dict(CASES[name])
        prob = synth.make_problem(c.pop("liks"), c.pop("N"), c.pop("M"), c.pop("Q"), Xdim=c.pop("Xdim"), seed=7, **c)
        for prec in ("fp64", "fp32", "tc"):
            t0 = time.time()
            try:
                err, out, o = pu.compare(prob, prec)
            except Exception:
                print("CASE %s %s FAILED" % (name, prec))
                traceback.print_exc()
                continue
            st = err.pop("_status")
            print("CASE %s %s  (%.1fs) elbo=%.10g oracle=%.10g status=%s" % (name, prec, time.time() - t0, out["log_marginal"][0, 0], o["log_marginal"][0, 0], st))
            print("   " + "  ".join("%s=%.2e" % (k, v) for k, v in err.items()))
            sys.stdout.flush()
