"""Dump the hand-off timestamps of the pair Gram kernel's first chunks (library built with -DHM_G2_TRACE=1)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from oracle import synth
import parity_util as pu
from hetmogp_b200._lib import lib
prob = synth.make_config("cfg3", N=int(sys.argv[1]) if len(sys.argv) > 1 else 200000)
eng = pu.make_engine(prob, "tc")
p = pu.params_of(prob)
for _ in range(3):
    eng.evaluate(p, what="full")
out = np.zeros((8, 2048), dtype=np.int64)
lib.hmogp_debug_g2_trace.restype = C.c_int
print("rc", lib.hmogp_debug_g2_trace(out.ctypes.data_as(C.c_void_p)))
names = ["gen: rowfull ok", "gen: empty ok", "gen: arrived", "iss: before wait", "iss: full ok", "iss: MMAs issued", "iss: committed"]
t0 = out[3, 0]
sl = slice(200, 232)
np.set_printoptions(linewidth=250)
for k in range(7):
    print("%-18s" % names[k], (out[k, sl] - t0))
d = lambda a: np.diff(out[a, 100:1500])
print("period per chunk: issuer %.0f  generator %.0f" % (d(4).mean(), d(1).mean()))
print("gen  empty-ok -> arrived      %.0f" % (out[2, 100:1500] - out[1, 100:1500]).mean())
print("gen  rowfull-ok -> empty-ok   %.0f" % (out[1, 100:1500] - out[0, 100:1500]).mean())
print("iss  arrived(gen) -> full ok  %.0f" % (out[4, 100:1500] - out[2, 100:1500]).mean())
print("iss  wait duration            %.0f" % (out[4, 100:1500] - out[3, 100:1500]).mean())
print("iss  full ok -> MMAs issued   %.0f" % (out[5, 100:1500] - out[4, 100:1500]).mean())
print("iss  commit                   %.0f" % (out[6, 100:1500] - out[5, 100:1500]).mean())
print("commit(c) -> gen empty ok(c+3) %.0f" % (out[1, 103:1503] - out[6, 100:1500]).mean())
