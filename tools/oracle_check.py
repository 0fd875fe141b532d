"""Engine (any mode) against the CPU oracle on a BASELINE configuration's shape at a bounded N: every block, the M x M
factors, KL / VE, per-row moments.   python tools/oracle_check.py <cfg> <N> [precision]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from oracle import synth  # noqa: E402
import parity_util as pu  # noqa: E402

cfg, N = sys.argv[1], int(sys.argv[2])
prec = sys.argv[3] if len(sys.argv) > 3 else "tc"
prob = synth.make_config(cfg, N=N)
t0 = time.time()
err, out, o = pu.compare(prob, prec, rows=True)
st = err.pop("_status")
print("ORACLE %s N=%d %s (%.1fs) elbo=%.10g oracle=%.10g status=%s" % (cfg, N, prec, time.time() - t0, out["log_marginal"][0, 0],
                                                                        o["log_marginal"][0, 0], st))
print("   " + "  ".join("%s=%.2e" % (k, v) for k, v in err.items()))
