"""CPU restatement of the reference's stochastic loop: util.vem_algorithm(stochastic=True) (hetmogp/util.py:320-329)
driving SVMOGP.stochastic_grad (hetmogp/svmogp.py:188-199) through paramz' optimizer_array / _grads.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  One iteration =
  new_batch (svmogp.py:175-186; slices of util.mini_slices served in order, quirk C-7; slice 0 went to the constructor)
  -> look-ahead of climin's Adadelta -> parameters_changed on the batch (oracle/diag_oracle.py, batch_scale =
  N_all / N_batch, svmogp.py:89-90) with the VE / VM gating of svmogp.py:104-166 -> transformed negative gradient ->
  Adadelta update -> VE/VM toggle (three... four VE steps, one VM step, svmogp.py:191-198).
The flat vector follows paramz' link order (svmogp.py:71-75): Z, m_u, L_u, [variance_q, lengthscale_q]..., [W_q,
kappa_q]...; ``fixed`` names the parameter groups left out ('Z', 'm_u', 'L_u', 'variance', 'lengthscale', 'W', 'kappa').
"""
import numpy as np

from . import climin_adadelta as ca
from . import diag_oracle


def mini_slices(n_samples, batch_size):
    """util.py:52-60."""
    n_batches, rest = divmod(n_samples, batch_size)
    if rest != 0:
        n_batches += 1
    return [slice(i * batch_size, (i + 1) * batch_size) for i in range(n_batches)]


def slice_stream(n_samples, batch_size):
    """util.py:62-72 with with_replacement=False: random.shuffle acts on a temporary list, so the order never changes."""
    slices = mini_slices(n_samples, batch_size)
    while True:
        for s in slices:
            yield s


def _blocks(prob, fixed):
    """(name, getter, setter, positive, variational, gradient-key, column) in link order."""
    Q = prob["Q"]
    out = []
    if "Z" not in fixed:
        out.append(("Z", None, False, False))
    if "m_u" not in fixed:
        out.append(("m_u", None, False, True))
    if "L_u" not in fixed:
        out.append(("L_u", None, False, True))
    for q in range(Q):
        if "variance" not in fixed:
            out.append(("rbf_var", q, True, False))
        if "lengthscale" not in fixed:
            out.append(("rbf_ls", q, True, False))
    for q in range(Q):
        if "W" not in fixed:
            out.append(("W", q, False, False))
        if "kappa" not in fixed:
            out.append(("kappa", q, True, False))
    return out


def get_flat(prob, fixed):
    parts = []
    for name, q, pos, _ in _blocks(prob, fixed):
        v = prob[name]
        v = v.ravel() if q is None else (np.atleast_1d(v[q]) if v.ndim == 1 else v[:, q])
        parts.append(ca.logexp_finv(v) if pos else np.array(v, dtype=np.float64))
    return np.concatenate(parts)


def set_flat(prob, fixed, x):
    i = 0
    for name, q, pos, _ in _blocks(prob, fixed):
        v = prob[name]
        if q is None:
            n = v.size
            v[...] = x[i:i + n].reshape(v.shape)
        elif v.ndim == 1:
            n = 1
            v[q] = ca.logexp_f(x[i:i + 1])[0] if pos else x[i]
        else:
            n = v.shape[0]
            v[:, q] = ca.logexp_f(x[i:i + n]) if pos else x[i:i + n]
        i += n


def flat_gradient(prob, fixed, o, ve_active, vm_active):
    """-(transformed gradient) exactly as paramz' _grads returns it after parameters_changed's gating."""
    GK = {"Z": "dZ", "m_u": "dL_dmu_u", "L_u": "dL_dL_u", "rbf_var": "d_rbf", "rbf_ls": "d_rbf", "W": "dW", "kappa": "dkappa"}
    parts = []
    for name, q, pos, variational in _blocks(prob, fixed):
        on = ve_active if variational else vm_active
        v = prob[name]
        if name in ("m_u", "L_u"):
            g = np.hstack(o[GK[name]]).ravel() if on else np.zeros(v.size)
        elif name == "Z":
            g = o["dZ"].ravel() if on else np.zeros(v.size)
        elif name == "rbf_var":
            g = np.array([o["d_rbf"][q, 0]]) if on else np.zeros(1)
        elif name == "rbf_ls":
            g = np.array([o["d_rbf"][q, 1]]) if on else np.zeros(1)
        else:
            g = o[GK[name]][:, q] if on else np.zeros(v.shape[0])
        if pos:
            th = np.atleast_1d(v[q]) if v.ndim == 1 else v[:, q]
            g = ca.logexp_gradfactor(th, g)
        parts.append(-g)
    return np.concatenate(parts)


def run(problem, batch_size, n_iters, step_rate=0.01, momentum=0.9, decay=0.9, offset=1e-4, fixed=("kappa", "lengthscale"),
        W_chain=None, kappa_chain=None):
    """Returns (elbo trace [n_iters], final problem dict at the last look-ahead point, optimiser state, last flat
    gradient).  ``problem`` is not modified."""
    prob = dict(problem)
    for k in ("Z", "m_u", "L_u", "rbf_var", "rbf_ls", "W", "kappa"):
        prob[k] = np.array(problem[k], dtype=np.float64)
    T = len(prob["Y"])
    N_all = [x.shape[0] for x in prob["X"]]
    streams = [slice_stream(n, batch_size) for n in N_all]
    for s in streams:
        next(s)                                                       # the constructor's batch (svmogp.py:46)
    wrt = get_flat(prob, fixed)
    st = ca.State(wrt.size, step_rate, decay, momentum, offset)
    vem_step, ve_count = True, 0
    trace = np.zeros(n_iters)
    g = None
    for it in range(n_iters):
        sl = [next(s) for s in streams]
        n_b = [len(range(*s.indices(n))) for s, n in zip(sl, N_all)]
        prob["batch_scale"] = [float(N_all[t] / n_b[t]) for t in range(T)]
        step1 = ca.lookahead(st, wrt)
        set_flat(prob, fixed, wrt)
        ve_active, vm_active = vem_step, (not vem_step)
        o = diag_oracle.elbo_and_grads(prob, row_slices=sl, W_chain=W_chain, kappa_chain=kappa_chain)
        trace[it] = o["log_marginal"][0, 0]
        g = flat_gradient(prob, fixed, o, ve_active, vm_active)
        ca.update(st, wrt, step1, g)
        if vem_step:                                                  # svmogp.py:191-198
            if ve_count > 2:
                ve_count, vem_step = 0, False
            else:
                ve_count += 1
        else:
            vem_step = True
    return trace, prob, st, g
