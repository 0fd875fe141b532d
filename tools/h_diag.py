"""Where does the tensor-core Gram lose accuracy?  Compare the packed statistics (g, H) of the tc mode with the fp64 mode
of the same engine, entry by entry: mean relative deviation (a uniform scaling of H is harmless), its scatter (what
K_uu^-1 . K_uu^-1 amplifies), by distance from the diagonal; and the errors of the blocks derived from H.

  python tools/h_diag.py <cfg> <N> [what]         env knobs of the engine apply (HMOGP_TC_FLUSH_ROWS, ...)
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from oracle import synth  # noqa: E402
import parity_util as pu  # noqa: E402
from hetmogp_b200 import _lib  # noqa: E402

GRADS = ("dL_dmu_u", "dL_dL_u", "dL_dKmm", "d_rbf", "dW", "dkappa", "dZ")


def stats_of(prob, prec, what=2):
    import torch
    eng = pu.make_engine(prob, prec)
    p = pu.params_of(prob)
    keep = []
    ps = eng._params(p, keep)
    n = int(_lib.lib.hmogp_stats_len(eng._h))
    buf = torch.zeros(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib.hmogp_step_local(eng._h, C.byref(ps), 0, what, C.c_void_p(buf.data_ptr())))
    torch.cuda.synchronize()
    out, gs = eng._alloc_out(what, False, True)
    st = _lib.Status()
    _lib.check(_lib.lib.hmogp_step_finish(eng._h, C.c_void_p(buf.data_ptr()), C.byref(gs), 0, what, C.byref(st)))
    out = {k: np.array(v) for k, v in out.items()}
    s = buf.cpu().numpy()
    eng.close()
    return s, out


def main():
    cfg, N = sys.argv[1], int(sys.argv[2])
    prob = synth.make_config(cfg, N=N)
    T, J, Q, M, Xd = prob["T"], prob["J"], prob["Q"], prob["M"], prob["Xdim"]
    Mp = 256
    while Mp < M:
        Mp *= 2
    off_g1 = 2 * T + J + 2 * J * Q + Q
    off_dz = off_g1 + Q * Mp
    off_H = (off_dz + Q * Xd * Mp + 1) & ~1
    s64, o64 = stats_of(prob, "fp64")
    stc, otc = stats_of(prob, "tc")
    print("ELBO rel %.3e" % (abs(otc["log_marginal"][0, 0] - o64["log_marginal"][0, 0]) / abs(o64["log_marginal"][0, 0])))
    print("BLOCKS " + "  ".join("%s=%.2e" % (k, pu.relerr(otc[k], o64[k])) for k in GRADS))
    for q in range(Q):
        H64 = s64[off_H + q * Mp * Mp: off_H + (q + 1) * Mp * Mp].reshape(Mp, Mp)[:M, :M]
        Htc = stc[off_H + q * Mp * Mp: off_H + (q + 1) * Mp * Mp].reshape(Mp, Mp)[:M, :M]
        g64, gtc = s64[off_g1 + q * Mp: off_g1 + q * Mp + M], stc[off_g1 + q * Mp: off_g1 + q * Mp + M]
        scale = np.abs(H64).max()
        rel = (Htc - H64) / np.where(np.abs(H64) > 1e-12 * scale, H64, np.inf)
        ii, jj = np.indices(H64.shape)
        print("q=%d  |H|max=%.3e  max|dH|/|H|max=%.3e  g relerr=%.3e" % (q, scale, np.abs(Htc - H64).max() / scale,
                                                                           np.abs(gtc - g64).max() / np.abs(g64).max()))
        for band in (0, 1, 2, 4, 8):
            m = (np.abs(ii - jj) == band) & (np.abs(H64) > 1e-12 * scale)
            r = rel[m]
            print("   |i-j|=%d  n=%d  mean rel dH=%+.3e  std=%.3e  max=%.3e   mean|H|/|H|max=%.2e" % (
                band, r.size, r.mean(), r.std(), np.abs(r).max(), np.abs(H64[m]).mean() / scale))
        # how much of the E error is the uniform part?
        Kuu_e = None
    # amplification: E = Ki H Ki from both H, with the fp64 K_uu^-1 (numpy)
    Z = prob["Z"]
    for q in range(Q):
        z = Z[:, q * Xd:(q + 1) * Xd]
        d2 = ((z[:, None, :] - z[None, :, :]) ** 2).sum(-1)
        Kuu = prob["rbf_var"][q] * np.exp(-0.5 * d2 / prob["rbf_ls"][q] ** 2)
        Ki = np.linalg.inv(Kuu)
        H64 = s64[off_H + q * Mp * Mp: off_H + (q + 1) * Mp * Mp].reshape(Mp, Mp)[:M, :M]
        Htc = stc[off_H + q * Mp * Mp: off_H + (q + 1) * Mp * Mp].reshape(Mp, Mp)[:M, :M]
        E64, Etc = Ki @ H64 @ Ki, Ki @ Htc @ Ki
        c = (Htc * H64).sum() / (H64 * H64).sum()          # best uniform scale
        Esc = Ki @ (Htc / c) @ Ki
        print("q=%d cond=%.2e  E relerr=%.3e   after removing the best uniform scale (c-1=%+.2e): %.3e" % (
            q, np.linalg.cond(Kuu), np.abs(Etc - E64).max() / np.abs(E64).max(), c - 1, np.abs(Esc - E64).max() / np.abs(E64).max()))


if __name__ == "__main__":
    main()
