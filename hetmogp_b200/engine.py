"""Host-side handle of the CUDA engine (one per process / GPU).

Holds the resident data shard of every task and the workspaces; each ``evaluate`` is one evaluation equivalent
to the reference's ``SVMOGP.parameters_changed()`` (hetmogp/svmogp.py:85-166).  With ``torch.distributed``
initialised and ``group`` given, rows are sharded across ranks and the per-shard sufficient statistics are summed
with ONE all-reduce per evaluation (NCCL on GPUs; gloo in the CPU tests of the host logic).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, f64, ptr


class Engine(object):
    print_v_negative = True      # the reference's stdout warning (svmogp_inf.py:221-222); bench.py keeps its stdout to one JSON line
    MAX_PINNED_SETS = 8   # page-locked result sets per output signature; further live sets use pageable memory

    def __init__(self, lik_specs, M, Q, Xdim, precision="fp32", device=0, group=None):
        self.lik_specs = [tuple(s) for s in lik_specs]
        self.T = len(self.lik_specs)
        self.M, self.Q, self.Xdim = int(M), int(Q), int(Xdim)
        self.P = self.M * (self.M + 1) // 2
        self.precision = precision
        self.device = int(device)
        self.group = group
        self._descs = (_lib.LikDesc * self.T)(*[_lib.lik_desc(s) for s in self.lik_specs])
        self.dimf = []
        for t in range(self.T):
            f = C.c_int32()
            check(lib.hmogp_lik_dims(C.byref(self._descs[t]), None, C.byref(f), None))
            self.dimf.append(f.value)
        self.J = int(sum(self.dimf))
        cfg = _lib.Config(self.M, self.Q, self.Xdim, self.T, _lib.PRECISIONS[precision], self.device, self._descs)
        self._h = C.c_void_p()
        check(lib.hmogp_create(C.byref(cfg), C.byref(self._h)))
        self.N = [0] * self.T
        self.status = None
        self._stats = None
        self._host_out = {}    # pools of page-locked result buffer sets, per output signature (see _alloc_out)

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.hmogp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream):
        check(lib.hmogp_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    # ------------------------------------------------------------------ data
    def set_data(self, X, Y):
        """X list[T] of (N_t, Xdim), Y list[T] of (N_t, 1) -- numpy (host) or torch CUDA float64 tensors."""
        self._keep = []
        for t in range(self.T):
            x, y = X[t], Y[t]
            if isinstance(x, np.ndarray):
                x, y = f64(x).reshape(-1, self.Xdim), f64(y).reshape(-1)
                kind = _lib.MEM_HOST
            else:
                x, y = x.contiguous().reshape(-1, self.Xdim), y.contiguous().reshape(-1)
                kind = _lib.MEM_DEVICE
            n = int(x.shape[0])
            assert int(y.shape[0]) == n, "X[%d] and Y[%d] disagree on N" % (t, t)
            check(lib.hmogp_set_data(self._h, t, ptr(x), ptr(y), n, kind))
            self.N[t] = n
            self._keep.append((x, y))   # pinned host buffers are uploaded by the next step (hmogp_set_data): keep them alive
        self._count = list(self.N)

    def set_rows(self, begin=None, count=None):
        if begin is None:
            check(lib.hmogp_set_rows(self._h, None, None))
            self._count = list(self.N)
            return
        b = (C.c_int64 * self.T)(*[int(v) for v in begin])
        c = (C.c_int64 * self.T)(*[int(v) for v in count])
        check(lib.hmogp_set_rows(self._h, b, c))
        self._count = [int(v) for v in count]

    def sync(self):
        import torch
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ evaluation
    def _params(self, p, keep):
        def g(name, shape=None):
            a = p.get(name)
            if a is None:
                return None
            if isinstance(a, np.ndarray) or not hasattr(a, "data_ptr"):
                a = f64(a)
            else:
                a = a.contiguous()
            keep.append(a)
            return ptr(a)
        ps = _lib.Params()
        for n in ("Z", "m_u", "L_u", "rbf_var", "rbf_ls", "W", "kappa", "W_chain", "kappa_chain", "batch_scale"):
            setattr(ps, n, g(n))
        return ps

    def _alloc_out(self, what, on_device, want_dKmm):
        M, Q, J, T, Xd, P = self.M, self.Q, self.J, self.T, self.Xdim, self.P
        shapes = {"log_marginal": (1, 1), "VE": (T,), "KL": (1,)}
        if what >= _lib.WHAT_VE:
            shapes.update(dL_dmu_u=(M, Q), dL_dL_u=(P, Q))
            if want_dKmm or what >= _lib.WHAT_FULL:
                shapes.update(dL_dKmm=(Q, M, M))
        if what >= _lib.WHAT_FULL:
            shapes.update(d_rbf=(Q, 2), dW=(J, Q), dkappa=(J, Q), dZ=(M, Q * Xd))
        import torch
        if on_device:
            out = {k: torch.empty(s, dtype=torch.float64, device="cuda:%d" % self.device) for k, s in shapes.items()}
        else:
            # Host results land in page-locked buffers (a device-to-host copy into pageable memory is staged by the driver
            # at a fifth of the speed: 1 ms of a 48 ms step for the 9 MB of cfg3).  The returned numpy arrays are views of
            # a pooled buffer set and own it for as long as any of them (or a slice of them) is referenced: a set is
            # handed out again only when the caller has dropped every array of it, so results never alias a later call's
            # (the reference returns fresh arrays, svmogp_inf.py:107-109).
            import sys
            key = (what, bool(want_dKmm))
            pool = self._host_out.setdefault(key, [])
            pinned = None
            for cand in pool:
                # a master array is free when nothing but the pool refers to it (every numpy view of it, and every
                # slice of a view, has it as .base): the dict entry, the loop variable, the argument of getrefcount
                if all(sys.getrefcount(m) <= 3 for m in cand.values()):
                    pinned = cand
                    break
            if pinned is None:
                pin = len(pool) < self.MAX_PINNED_SETS
                pinned = {k: torch.empty(s, dtype=torch.float64, pin_memory=pin).numpy() for k, s in shapes.items()}
                pool.append(pinned)
            out = {k: m.view() for k, m in pinned.items()}
        gs = _lib.Grads()
        for k, a in out.items():
            setattr(gs, k, ptr(a))
        return out, gs

    def evaluate(self, params, what="full", want_dKmm=False, out=None, hyper_unchanged=False):
        """params: dict with Z, m_u, L_u, rbf_var, rbf_ls, W, kappa [, W_chain, kappa_chain, batch_scale] as numpy
        arrays (host path: copies inside the call) or torch CUDA tensors (device path).  Returns a dict of outputs
        in the reference's layouts (see include/hetmogp_b200.h); ``self.status`` holds the flags.  With host parameters the
        returned numpy arrays are backed by pooled page-locked buffers that stay theirs for as long as they are referenced
        (no aliasing between calls).  The factorisation of K_uu is reused between calls whose Z / rbf_var / rbf_ls are
        bitwise equal (host parameters: detected; CUDA-tensor parameters: only with ``hyper_unchanged=True``)."""
        if hyper_unchanged:
            check(lib.hmogp_hint_hyper_unchanged(self._h, 1))
        w = {"elbo": _lib.WHAT_ELBO, "ve": _lib.WHAT_VE, "full": _lib.WHAT_FULL}[what] if isinstance(what, str) else what
        on_device = hasattr(params["m_u"], "data_ptr")
        kind = _lib.MEM_DEVICE if on_device else _lib.MEM_HOST
        keep = []
        ps = self._params(params, keep)
        if out is None:
            out, gs = self._alloc_out(w, on_device, want_dKmm)
        else:
            gs = _lib.Grads()
            for k, a in out.items():
                setattr(gs, k, ptr(a))
        st = _lib.Status()
        if self.group is None:
            check(lib.hmogp_elbo_and_grads(self._h, C.byref(ps), C.byref(gs), kind, w, C.byref(st)))
        else:
            import torch
            import torch.distributed as dist
            if self._stats is None:
                n = int(lib.hmogp_stats_len(self._h))
                self._stats = torch.empty(n, dtype=torch.float64, device="cuda:%d" % self.device)
            # the all-reduce runs on torch's current stream: keep the engine on the same one, every call
            self.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
            check(lib.hmogp_step_local(self._h, C.byref(ps), kind, w, C.c_void_p(self._stats.data_ptr())))
            dist.all_reduce(self._stats, op=dist.ReduceOp.SUM, group=self.group)   # the ONE collective of a step
            check(lib.hmogp_step_finish(self._h, C.c_void_p(self._stats.data_ptr()), C.byref(gs), kind, w, C.byref(st)))
        self.status = {"jitter": [st.jitter[q] for q in range(self.Q)],
                       "chol_fail": [st.chol_fail[q] for q in range(self.Q)],
                       "lu_singular": [st.lu_singular[q] for q in range(self.Q)],
                       "n_negative_v": int(st.n_negative_v)}
        if self.status["n_negative_v"] > 0 and Engine.print_v_negative:
            print('v negative!')   # svmogp_inf.py:221-222 (warning only)
        return out

    @property
    def kuu_reuse_count(self):
        return int(lib.hmogp_kuu_reuse_count(self._h))

    def predict_f(self, params, t, Xnew):
        """q(f_d) at new inputs for the output functions of task t: (m_fd, v_fd), each (N, dim_f[t]) -- the forward
        projection and the W-mix only, no labels, no likelihood (hmogp_predict_f)."""
        keep = []
        ps = self._params(params, keep)
        if isinstance(Xnew, np.ndarray) or not hasattr(Xnew, "data_ptr"):
            x = f64(Xnew).reshape(-1, self.Xdim)
            n = int(x.shape[0])
            m, v = np.empty((n, self.dimf[t])), np.empty((n, self.dimf[t]))
            kind = _lib.MEM_HOST
        else:
            import torch
            x = Xnew.contiguous().reshape(-1, self.Xdim)
            n = int(x.shape[0])
            m = torch.empty((n, self.dimf[t]), dtype=torch.float64, device=x.device)
            v = torch.empty_like(m)
            kind = _lib.MEM_DEVICE
        if (kind == _lib.MEM_DEVICE) != hasattr(params["m_u"], "data_ptr"):
            raise ValueError("predict_f: parameters and Xnew must both be host arrays or both CUDA tensors")
        check(lib.hmogp_predict_f(self._h, C.byref(ps), kind, int(t), ptr(x), n, ptr(m), ptr(v)))
        return m, v

    # ------------------------------------------------------------------ introspection (tests, small N)
    def rows(self, t):
        n = self._active_count(t)
        F = self.dimf[t]
        m, v, dm, dv = (np.empty((n, F)) for _ in range(4))
        ve = np.empty((n, 1))
        check(lib.hmogp_get_rows(self._h, t, ptr(m), ptr(v), ptr(ve), ptr(dm), ptr(dv)))
        return {"m": m, "v": v, "ve": ve, "dm": dm, "dv": dv}

    def _active_count(self, t):
        return getattr(self, "_count", self.N)[t]

    def dense_dL_dKmn(self, q, d):
        t = int(np.searchsorted(np.cumsum(self.dimf), d, side="right"))
        n = self._active_count(t)
        out = np.empty((self.M, n))
        diag = np.empty(n)
        check(lib.hmogp_get_dL_dKmn(self._h, q, d, ptr(out), ptr(diag)))
        return out, diag

    def kuu(self):
        Q, M = self.Q, self.M
        Kuu, Luu, Kuui = np.empty((Q, M, M)), np.empty((Q, M, M)), np.empty((Q, M, M))
        check(lib.hmogp_get_kuu(self._h, ptr(Kuu), ptr(Luu), ptr(Kuui)))
        return Kuu, Luu, Kuui

    def enable_timing(self, on=True):
        check(lib.hmogp_enable_timing(self._h, 1 if on else 0))

    def last_timing(self):
        f = (C.c_float * 6)()
        n = C.c_int32()
        check(lib.hmogp_last_timing(self._h, f, C.byref(n)))
        names = ("prepare_ms", "forward_ms", "lik_ms", "bwd_proj_ms", "bwd_gram_ms", "finish_ms")
        d = {k: float(f[i]) for i, k in enumerate(names)}
        d["launches"] = n.value
        return d


def shard_rows(N, rank, world):
    """Contiguous row shard [begin, begin+count) of rank `rank` out of `world`."""
    begin = [(n * rank) // world for n in N]
    end = [(n * (rank + 1)) // world for n in N]
    return begin, [e - b for b, e in zip(begin, end)]
