"""Device-resident stochastic variational inference loop (reference: util.vem_algorithm(stochastic=True),
/root/reference/hetmogp/util.py:320-329, driving SVMOGP.stochastic_grad, svmogp.py:188-199).

Everything an iteration touches stays in HBM: the parameter arrays in the engine's layouts (hmogp_params), the gradient
arrays the engine writes (hmogp_grads) and the Adadelta state over paramz' flat optimizer vector (csrc/optim.cu).  Per
iteration the host only advances the minibatch slice (two integers per task) and reads nothing back; the ELBO trace is
copied once at the end.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check
from .svmogp_inf import flatten_params


class DeviceSVI(object):
    def __init__(self, model, step_rate, momentum, decay, offset):
        import torch
        self.model = model
        eng = model._eng
        self.dev = torch.device("cuda", eng.device)
        self.hyper = (float(step_rate), float(momentum), float(decay), float(offset))
        self.fix_pattern = self._pattern(model)
        Q, J = model.num_latent_funcs, model.num_output_funcs
        p = flatten_params(model.q_u_means, model.q_u_chols, model.Z, model.kern_list, model.B_list, None,
                           model._W0 if model.compat_stale_W else None, model._k0 if model.compat_stale_W else None)
        self.params = {k: torch.as_tensor(v, device=self.dev).contiguous() for k, v in p.items()}
        self._bs_cache = {}
        self.out, _ = eng._alloc_out(_lib.WHAT_FULL, True, False)
        for v in self.out.values():
            v.zero_()
        self.out_ve = {k: self.out[k] for k in ("log_marginal", "VE", "KL", "dL_dmu_u", "dL_dL_u")}
        # ---- segment table in paramz link order (svmogp.py:71-75), unfixed parameters only
        P, G = self.params, self.out
        segs = []

        def add(param_t, grad_t, count, stride, positive, variational, p_off=0, g_off=0):
            segs.append((param_t.data_ptr() + 8 * p_off, grad_t.data_ptr() + 8 * g_off, int(count), int(stride), int(positive), int(variational)))

        if not model.Z.is_fixed:
            add(P["Z"], G["dZ"], P["Z"].numel(), 1, 0, 0)
        if not model.q_u_means.is_fixed:
            add(P["m_u"], G["dL_dmu_u"], P["m_u"].numel(), 1, 0, 1)
        if not model.q_u_chols.is_fixed:
            add(P["L_u"], G["dL_dL_u"], P["L_u"].numel(), 1, 0, 1)
        for q, k in enumerate(model.kern_list):
            if not k.variance.is_fixed:
                add(P["rbf_var"], G["d_rbf"], 1, 1, 1, 0, p_off=q, g_off=2 * q)
            if not k.lengthscale.is_fixed:
                add(P["rbf_ls"], G["d_rbf"], 1, 1, 1, 0, p_off=q, g_off=2 * q + 1)
        for q, B in enumerate(model.B_list):
            if not B.W.is_fixed:
                add(P["W"], G["dW"], J, Q, 0, 0, p_off=q, g_off=q)
            if not B.kappa.is_fixed:
                add(P["kappa"], G["dkappa"], J, Q, 1, 0, p_off=q, g_off=q)
        if not segs:
            raise ValueError("svi_device: every parameter is fixed")
        arr = (_lib.OptSegment * len(segs))()
        off = 0
        for i, (pp, gp, n, stride, pos, var) in enumerate(segs):
            arr[i].offset, arr[i].count, arr[i].param, arr[i].grad = off, n, pp, gp
            arr[i].stride, arr[i].positive, arr[i].variational, arr[i].reserved = stride, pos, var, 0
            off += n
        self.n = off
        self._h = C.c_void_p()
        check(lib.hmogp_opt_create(eng.device, arr, len(segs), step_rate, decay, momentum, offset, C.byref(self._h)))
        self.stream = torch.cuda.current_stream(self.dev).cuda_stream
        check(lib.hmogp_opt_gather(self._h, C.c_void_p(self.stream)))

    @staticmethod
    def _pattern(model):
        return tuple(bool(p.is_fixed) for _, p, _, _ in model._named())

    def matches(self, model, step_rate, momentum, decay, offset):
        return model is self.model and self.hyper == (float(step_rate), float(momentum), float(decay), float(offset)) and \
            self.fix_pattern == self._pattern(model)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib.hmogp_opt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def state(self, which):
        """Host copy of a state vector: 'wrt' (the flat optimizer_array), 'gms', 'sms' or 'step'."""
        out = np.empty(self.n)
        check(lib.hmogp_opt_get_state(self._h, {"wrt": 0, "gms": 1, "sms": 2, "step": 3}[which], out.ctypes.data,
                                      C.c_void_p(self.stream)))
        return out

    def _batch_scale_tensor(self, scales):
        import torch
        key = tuple(scales)
        t = self._bs_cache.get(key)
        if t is None:
            t = torch.tensor(list(key), dtype=torch.float64, device=self.dev)
            self._bs_cache[key] = t
        return t

    def run(self, n_iters, trace=True):
        import torch
        m, eng = self.model, self.model._eng
        if not m.stochastic:
            raise ValueError("svi_device needs a model built with batch_size (svmogp.py:38-47)")
        eng.set_stream(self.stream)
        tr = torch.zeros(max(1, n_iters), dtype=torch.float64, device=self.dev)
        s = C.c_void_p(self.stream)
        ve_active = vm_active = True
        for it in range(n_iters):
            m.new_batch()                                                   # svmogp.py:189
            eng.set_rows(*m._rank_rows())
            self.params["batch_scale"] = self._batch_scale_tensor(m._batch_scale())
            check(lib.hmogp_opt_lookahead(self._h, 1, s))                   # wrt -= momentum * step; parameters <- wrt
            ve_active, vm_active = bool(m.vem_step), not bool(m.vem_step)   # svmogp.py:104-166 gating
            if vm_active:
                eng.evaluate(self.params, what="full", out=self.out)
            else:
                eng.evaluate(self.params, what="ve", out=self.out_ve)
            check(lib.hmogp_opt_update(self._h, int(ve_active), int(vm_active), None, s))
            if trace:
                tr[it:it + 1].copy_(self.out["log_marginal"].reshape(-1)[:1], non_blocking=True)
            m._advance_vem()                                                # svmogp.py:191-198
        self._sync_back(ve_active, vm_active)
        return tr[:n_iters].cpu().numpy()

    def _sync_back(self, ve_active, vm_active):
        """Leave the host-side Param objects as the reference's model is after minimize_until: parameters of the last
        look-ahead point, gradient fields as its parameters_changed wrote them."""
        m = self.model
        P = {k: v.cpu().numpy() for k, v in self.params.items()}
        G = {k: v.cpu().numpy() for k, v in self.out.items()}
        Q = m.num_latent_funcs
        np.asarray(m.Z)[...] = P["Z"]
        np.asarray(m.q_u_means)[...] = P["m_u"]
        np.asarray(m.q_u_chols)[...] = P["L_u"]
        for q in range(Q):
            np.asarray(m.kern_list[q].variance)[...] = P["rbf_var"][q]
            np.asarray(m.kern_list[q].lengthscale)[...] = P["rbf_ls"][q]
            np.asarray(m.B_list[q].W)[...] = P["W"][:, q:q + 1]
            np.asarray(m.B_list[q].kappa)[...] = P["kappa"][:, q]
        m._log_marginal_likelihood = G["log_marginal"].reshape(1, 1).copy()
        m.q_u_means.gradient = G["dL_dmu_u"].copy() if ve_active else np.zeros(m.q_u_means.shape)
        m.q_u_chols.gradient = G["dL_dL_u"].copy() if ve_active else np.zeros(m.q_u_chols.shape)
        for q in range(Q):
            if vm_active:
                m.kern_list[q].gradient = G["d_rbf"][q].copy()
                m.B_list[q].gradient = np.concatenate([G["dW"][:, q], G["dkappa"][:, q]])
            else:
                m.kern_list[q].gradient = np.zeros(2)
                m.B_list[q].gradient = np.zeros(2 * m.num_output_funcs)
        m.Z.gradient = G["dZ"].copy() if (vm_active and not m.Z.is_fixed) else np.zeros(m.Z.shape)
        m.status = m._eng.status
