"""How is the tensor-core Gram H^1 wrong?  Elementwise comparison with the fp64 mode's H^1 on the benchmark problem:
uniform relative bias (truncating accumulation) vs noise.   python tools/h_bias.py [N]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from oracle import synth
import parity_util as pu
from hetmogp_b200 import _lib
from hetmogp_b200._lib import lib, check
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
prob = synth.make_config("cfg3", N=N)
p = pu.params_of(prob)
M, Q = prob["M"], prob["Q"]
H = {}
for prec in ("fp64", "tc"):
    eng = pu.make_engine(prob, prec)
    n = int(lib.hmogp_stats_len(eng._h))
    st = torch.zeros(n, dtype=torch.float64, device="cuda")
    keep = []
    ps = eng._params(p, keep)
    check(lib.hmogp_step_local(eng._h, C.byref(ps), _lib.MEM_HOST, _lib.WHAT_FULL, C.c_void_p(st.data_ptr())))
    torch.cuda.synchronize()
    Mp = 512
    H[prec] = st[-Q * Mp * Mp:].reshape(Q, Mp, Mp)[:, :M, :M].cpu().numpy().copy()
    eng.close()
for q in range(Q):
    a, b = H["tc"][q], H["fp64"][q]
    big = np.abs(b) > 1e-3 * np.abs(b).max()
    r = (a[big] - b[big]) / b[big]
    d = np.diag(a) / np.diag(b) - 1
    print("q=%d  entries>1e-3 max: n=%d  rel err mean %.3e  std %.3e  median %.3e | diag mean %.3e std %.3e | global fit beta=%.3e resid rms/|H| %.3e"
          % (q, big.sum(), r.mean(), r.std(), np.median(r), d.mean(), d.std(), 1 - (a * b).sum() / (b * b).sum(),
             np.sqrt(((a - b * (a * b).sum() / (b * b).sum()) ** 2).mean()) / np.abs(b).max()))
    for k in (0, 1, 2, 4, 8):
        dk = np.diagonal(a, k) / np.diagonal(b, k) - 1
        print("      off-diagonal %d: mean %.3e std %.3e" % (k, dk.mean(), dk.std()))
