"""SVMOGPInf with the reference's interface (hetmogp/svmogp_inf.py:21-109), backed by the CUDA engine.

``inference(q_u_means, q_u_chols, X, Y, Z, kern_list, likelihood, B_list, Y_metadata, KL_scale=1.0,
batch_scale=None, predictive=False)`` returns ``(log_marginal, gradients, posteriors, posteriors_F)`` like the
reference (svmogp_inf.py:109).  Differences, all of them about N-sized objects the reference materialises and
training never reads:
  * ``gradients['dL_dKmn']`` / ``['dL_dKdiag']`` are lazy: element ``[q][d]`` is computed on the GPU on access
    (dense (M, N_t) / (N_t,), svmogp_inf.py:157-164) -- meant for small N;
  * ``posteriors_F`` is a lazy list: element ``[d]`` carries the marginal mean / variance of q(f_d) at the training
    inputs (svmogp_inf.py:216-218), computed on access, not an N x N GPy Posterior (svmogp_inf.py:48-50);
  * with ``predictive=True`` the call returns that list directly, as the reference does (svmogp_inf.py:52), evaluated at
    the given ``X`` -- whose row counts may differ from ``Y``'s (svmogp.py:291): only X is uploaded, no likelihood runs;
  * the dict additionally holds the hyper-parameter chain rule of svmogp.py:100-166 computed by the engine
    (``d_rbf``, ``dW``, ``dkappa``, ``dZ``) so ``SVMOGP.parameters_changed`` never needs the dense blocks.
``KL_scale`` is accepted and ignored exactly as in the reference (svmogp_inf.py:23, quirk C-6).

Like the reference, the call is stateless towards its caller: X and Y are uploaded on every call and the returned arrays
are the caller's (no later call overwrites them).  ``data_token`` opts into residency: pass any hashable value that changes
whenever the contents of X / Y change, and the upload is skipped while it stays the same.
"""
import collections

import numpy as np

from .engine import Engine

qfd = collections.namedtuple("q_fd", "m_fd, v_fd")                 # svmogp_inf.py:17 (diagonal part only)
PosteriorU = collections.namedtuple("PosteriorU", "mean, chol_flat")
PosteriorF = collections.namedtuple("PosteriorF", "mean, variance")


class _LazyBlocks(object):
    """gradients['dL_dKmn'][q][d] / ['dL_dKdiag'][q][d] computed on first access (small-N use)."""

    def __init__(self, eng, Q, J, which):
        self._eng, self._Q, self._J, self._which, self._cache = eng, Q, J, which, {}

    def __len__(self):
        return self._Q

    def __getitem__(self, q):
        outer = self

        class _Row(object):
            def __len__(self_inner):
                return outer._J

            def __getitem__(self_inner, d):
                key = (q, d)
                if key not in outer._cache:
                    outer._cache[key] = outer._eng.dense_dL_dKmn(q, d)
                return outer._cache[key][outer._which]
        return _Row()


class _LazyPosteriorF(object):
    """posteriors_F[d] = PosteriorF(mean (N_t,1), variance (N_t,1)) of q(f_d) at the rows of the last evaluation."""

    def __init__(self, eng, f_index, d_index):
        self._eng, self._f, self._d, self._rows = eng, f_index, d_index, {}

    def __len__(self):
        return len(self._f)

    def __getitem__(self, d):
        t = int(self._f[d])
        if t not in self._rows:
            self._rows[t] = self._eng.rows(t)
        k = int(self._d[d])
        return PosteriorF(self._rows[t]["m"][:, k:k + 1].copy(), self._rows[t]["v"][:, k:k + 1].copy())


def flatten_params(q_u_means, q_u_chols, Z, kern_list, B_list, batch_scale=None, W_chain=None, kappa_chain=None):
    """kern_list / B_list objects (GPy or gpy_shim) -> the flat arrays of hmogp_params."""
    Q = len(kern_list)
    p = dict(
        Z=np.ascontiguousarray(Z, dtype=np.float64),
        m_u=np.ascontiguousarray(q_u_means, dtype=np.float64),
        L_u=np.ascontiguousarray(q_u_chols, dtype=np.float64),
        rbf_var=np.array([float(np.asarray(k.variance).ravel()[0]) for k in kern_list]),
        rbf_ls=np.array([float(np.asarray(k.lengthscale).ravel()[0]) for k in kern_list]),
        W=np.ascontiguousarray(np.hstack([np.asarray(B.W, dtype=np.float64).reshape(-1, 1) for B in B_list])),
        kappa=np.ascontiguousarray(np.stack([np.asarray(B.kappa, dtype=np.float64).ravel() for B in B_list], axis=1)),
    )
    if batch_scale is not None:
        p["batch_scale"] = np.asarray(batch_scale, dtype=np.float64)
    if W_chain is not None:
        p["W_chain"] = np.ascontiguousarray(W_chain, dtype=np.float64)
    if kappa_chain is not None:
        p["kappa_chain"] = np.ascontiguousarray(kappa_chain, dtype=np.float64)
    assert p["W"].shape[1] == Q
    return p


class SVMOGPInf(object):
    def __init__(self, precision="fp32", device=0, group=None):
        self.precision, self.device, self.group = precision, device, group
        self._eng = None
        self._key = None
        self._data_token = None
        self.status = None

    def engine_for(self, likelihood, M, Q, Xdim):
        specs = tuple(tuple(l.spec) for l in likelihood.likelihoods_list)
        key = (specs, M, Q, Xdim, self.precision, self.device)
        if self._key != key:
            if self._eng is not None:
                self._eng.close()
            self._eng = Engine(specs, M, Q, Xdim, precision=self.precision, device=self.device, group=self.group)
            self._key, self._data_token = key, None
        return self._eng

    def inference(self, q_u_means, q_u_chols, X, Y, Z, kern_list, likelihood, B_list, Y_metadata, KL_scale=1.0,
                  batch_scale=None, predictive=False, what="full", W_chain=None, kappa_chain=None, data_token=None,
                  copy=False):
        """``what`` ('full' | 've' | 'elbo'), ``W_chain`` / ``kappa_chain`` (quirk C-5), ``data_token`` (see the module
        docstring) and ``copy`` are extensions.  The returned arrays are backed by pooled page-locked buffers that belong to
        the caller for as long as it references them (Engine._alloc_out); ``copy=True`` detaches them into ordinary numpy
        memory."""
        M, Q = np.asarray(q_u_means).shape[0], len(kern_list)
        Xdim = int(np.asarray(Z).shape[1] // Q)
        eng = self.engine_for(likelihood, M, Q, Xdim)
        params = flatten_params(q_u_means, q_u_chols, Z, kern_list, B_list, batch_scale, W_chain, kappa_chain)
        J = params["W"].shape[0]
        f_index, d_index = Y_metadata['function_index'].flatten(), Y_metadata['d_index'].flatten()
        if predictive:
            # svmogp_inf.py:43-52 with predictive=True: q(f_d) at X[function_index[d]]; Y is not read
            per_task = {}
            out = []
            for d in range(J):
                t = int(f_index[d])
                if t not in per_task:
                    per_task[t] = eng.predict_f(params, t, X[t])
                m, v = per_task[t]
                k = int(d_index[d])
                out.append(PosteriorF(m[:, k:k + 1].copy(), v[:, k:k + 1].copy()))
            return out
        if data_token is None or data_token != self._data_token:
            eng.set_data(X, Y)
            self._data_token = data_token
        eng.set_rows(None)                                # the call evaluates every row it is given
        out = eng.evaluate(params, what=what, want_dKmm=True)
        if copy:
            out = {k: np.array(v) for k, v in out.items()}
        log_marginal = out["log_marginal"]                                    # (1,1) like svmogp_inf.py:246
        gradients = {}
        if "dL_dmu_u" in out:
            gradients['dL_dmu_u'] = [out["dL_dmu_u"][:, q:q + 1] for q in range(Q)]
            gradients['dL_dL_u'] = [out["dL_dL_u"][:, q:q + 1] for q in range(Q)]
            gradients['dL_dKmm'] = [out["dL_dKmm"][q] for q in range(Q)]
            gradients['dL_dKmn'] = _LazyBlocks(eng, Q, J, 0)
            gradients['dL_dKdiag'] = _LazyBlocks(eng, Q, J, 1)
        for k in ("d_rbf", "dW", "dkappa", "dZ", "VE", "KL"):
            if k in out:
                gradients[k] = out[k]
        posteriors = [PosteriorU(params["m_u"][:, q:q + 1].copy(), params["L_u"][:, q:q + 1].copy()) for q in range(Q)]
        self.status = eng.status
        return log_marginal, gradients, posteriors, _LazyPosteriorF(eng, f_index, d_index)
