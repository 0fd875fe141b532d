"""SVMOGP model container with the reference's surface (hetmogp/svmogp.py:16-217), backed by the CUDA engine.

Keeps ``SVMOGP(X, Y, Z, kern_list, likelihood, Y_metadata, name, batch_size, W_list)``, ``parameters_changed()``
(writes the ``.gradient`` fields exactly as svmogp.py:100-166 leaves them, including the VE/VM gating of the
stochastic mode), ``log_likelihood()`` (a (1,1) array, svmogp.py:82-83), ``new_batch`` / ``set_data`` /
``stochastic_grad`` / ``callback`` (svmogp.py:168-217), a paramz-style flat ``optimizer_array`` with the Logexp
transform of positive parameters, regular-expression parameter selection (``model['.*.lengthscale'].fix()``, as
util.vem_algorithm uses it, util.py:285-318), ``optimize`` (L-BFGS-B like paramz' default) and the prediction entry
points ``_raw_predict`` / ``_raw_predict_f`` / ``predictive_new`` / ``predictive`` / ``negative_log_predictive``
(svmogp.py:219-370) on the O(M^2)-per-point route through q(U).

The data stay resident on the GPU; a minibatch is a row slice of it.  ``svi_device`` runs the stochastic loop of
util.vem_algorithm (climin Adadelta, util.py:320-329) without leaving the device: parameters, gradients and optimiser
state live in HBM and each iteration is one engine evaluation plus two small kernels (csrc/optim.cu).

Deviation from the reference, stated: the chain-rule multipliers of svmogp.py:141,143,156 are taken from the CURRENT W /
kappa by default; the reference rebuilds them from the constructor-time ``W_list`` (svmogp.py:98-99, quirk C-5 of
SURVEY.md).  ``compat_stale_W=True`` reproduces the reference.
"""
import ctypes as C
import re

import numpy as np

from . import _lib, util
from ._lib import lib, check
from .engine import Engine, shard_rows
from .gpy_shim import Param
from .svmogp_inf import SVMOGPInf, flatten_params

_LIM = 36.0                       # paramz transformations._lim_val
_LOG_LIM = 709.782712893384       # log(DBL_MAX): paramz clips the argument of exp


def _logexp_f(x):      # paramz Logexp.f: theta = log(1 + e^x)
    x = np.asarray(x, dtype=np.float64)
    return np.where(x > _LIM, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM, _LIM))))


def _logexp_finv(t):   # paramz Logexp.finv
    t = np.asarray(t, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return np.where(t > _LIM, t, np.log(np.expm1(np.minimum(t, _LIM))))


def _logexp_gradfactor(t):   # paramz Logexp.gradfactor / df
    t = np.asarray(t, dtype=np.float64)
    return np.where(t > _LIM, 1.0, -np.expm1(-t))


class _ParamGroup(object):
    """What ``model['regex']`` returns: the matching parameters, with fix() / unfix() (paramz indexing by name)."""

    def __init__(self, params):
        self.params = params

    def fix(self):
        for p in self.params:
            p.fix()

    def unfix(self):
        for p in self.params:
            p.unfix()

    @property
    def values(self):
        return [np.asarray(p) for p in self.params]

    def __len__(self):
        return len(self.params)


class SVMOGP(object):
    def __init__(self, X, Y, Z, kern_list, likelihood, Y_metadata, name='SVMOGP', batch_size=None, W_list=None,
                 precision="fp32", device=0, group=None, presharded=False, compat_stale_W=False):
        self.name = name
        self.batch_size = batch_size
        self.kern_list = kern_list
        self.likelihood = likelihood
        self.Y_metadata = Y_metadata
        self.num_inducing = Z.shape[0]
        self.num_latent_funcs = len(kern_list)
        self.num_output_funcs = likelihood.num_output_functions(self.Y_metadata)
        if W_list is None:
            self.W_list, self.kappa_list = util.random_W_kappas(self.num_latent_funcs, self.num_output_funcs, rank=1)
        else:
            self.W_list = W_list
            _, self.kappa_list = util.random_W_kappas(self.num_latent_funcs, self.num_output_funcs, rank=1)
        self.Xmulti_all, self.Ymulti_all = X, Y
        self.Xdim = Z.shape[1]
        self.Z = Param('inducing inputs', np.tile(Z, (1, self.num_latent_funcs)))          # svmogp.py:52
        self.inference_method = SVMOGPInf(precision=precision, device=device, group=group)
        _, self.B_list = util.LCM(input_dim=self.Xdim, output_dim=self.num_output_funcs, rank=1,
                                  kernels_list=self.kern_list, W_list=self.W_list, kappa_list=self.kappa_list)
        # stale chain multipliers of svmogp.py:98-99,141,143,156 (quirk C-5): constructor-time W, kappa
        self.compat_stale_W = compat_stale_W
        self._W0 = np.hstack([np.asarray(w, dtype=np.float64).reshape(-1, 1) for w in self.W_list])
        self._k0 = np.stack([np.asarray(k, dtype=np.float64).ravel() for k in self.kappa_list], axis=1)
        self.q_u_means = Param('m_u', 2.5 * np.random.randn(self.num_inducing, self.num_latent_funcs))   # svmogp.py:66
        M = self.num_inducing
        ii, jj = np.tril_indices(M)
        chols = np.tile(np.eye(M)[ii, jj][:, None], (1, self.num_latent_funcs))                            # svmogp.py:68
        self.q_u_chols = Param('L_u', chols)
        # The model owns its engine (the full data resident; minibatches are row slices, util.py:52-72).  With a process
        # group every rank holds the same rows and evaluates its share of the active slice; the per-rank statistics are
        # summed by the engine's one all-reduce.  presharded=True: the caller already gave this rank only its rows (the
        # batch scales then need the global row counts: pass them as N_global).
        specs = tuple(tuple(l.spec) for l in likelihood.likelihoods_list)
        self.group, self.presharded = group, presharded
        self._rank, self._world = 0, 1
        if group is not None:
            import torch.distributed as dist
            self._rank, self._world = dist.get_rank(group), dist.get_world_size(group)
        self._eng = Engine(specs, M, self.num_latent_funcs, self.Xdim, precision=precision, device=device, group=group)
        self._eng.set_data(X, Y)
        self._N_all = [int(x.shape[0]) for x in X]
        self.N_global = list(self._N_all)
        if presharded and group is not None:
            import torch
            import torch.distributed as dist
            n = torch.tensor(self._N_all, dtype=torch.int64, device="cuda:%d" % device)
            dist.all_reduce(n, group=group)
            self.N_global = [int(v) for v in n.cpu()]
        if batch_size is None:
            self.stochastic = False
            self._slice = [(0, n) for n in self._N_all]
        else:
            self.stochastic = True
            self.slicer_list = [util.draw_mini_slices(n, self.batch_size) for n in self._N_all]
            self.new_batch()                                                                   # svmogp.py:46 (slice 0)
        self.vem_step = True
        self.ve_count = 0
        self.elbo = np.zeros((1, 1))
        self._log_marginal_likelihood = np.zeros((1, 1))
        self.posteriors = None
        self._svi = None
        self.parameters_changed()

    # ------------------------------------------------------------------ data / minibatching (svmogp.py:168-186)
    @staticmethod
    def _clip(x, b, c):
        return x[b:b + c]

    @property
    def Xmulti(self):
        return [self._clip(x, b, c) for x, (b, c) in zip(self.Xmulti_all, self._slice)]

    @property
    def Ymulti(self):
        return [self._clip(y, b, c) for y, (b, c) in zip(self.Ymulti_all, self._slice)]

    def new_batch(self):
        """svmogp.py:175-186: the next slice of every task's slicer; the batch is data[slice] (numpy clamps a slice that
        runs past the end, so the last batch of an epoch may be short and its batch scale larger, svmogp.py:89-90)."""
        sl = [next(s) for s in self.slicer_list]
        self._slice = [(s.start, len(range(*s.indices(n)))) for s, n in zip(sl, self._N_all)]
        return self.Xmulti, self.Ymulti

    def set_data(self, X=None, Y=None):
        """svmogp.py:168-173.  The batch of this model is a slice of the resident data (``new_batch`` has already moved
        it); arrays passed here are accepted for interface parity and must be that batch."""
        if X is not None:
            for t, x in enumerate(X):
                if int(x.shape[0]) != self._slice[t][1]:
                    raise ValueError("set_data: task %d has %d rows, the current batch has %d" % (t, x.shape[0], self._slice[t][1]))

    def _rank_rows(self):
        """Rows [begin, begin+count) of every task this rank evaluates for the current slice."""
        if self._world == 1 or self.presharded:
            return [b for b, _ in self._slice], [c for _, c in self._slice]
        b0 = [b for b, _ in self._slice]
        sb, sc = shard_rows([c for _, c in self._slice], self._rank, self._world)
        return [b + s for b, s in zip(b0, sb)], sc

    # ------------------------------------------------------------------ the hot path (svmogp.py:82-166)
    def log_likelihood(self):
        return self._log_marginal_likelihood

    def _batch_scale(self):
        T = len(self.likelihood.likelihoods_list)
        if self.presharded and self._world > 1 and self.stochastic:
            raise NotImplementedError("minibatching over presharded data")
        return [float(self.N_global[t] / self._slice[t][1]) if not (self.presharded and self._world > 1)
                else 1.0 for t in range(T)]                                                     # svmogp.py:89-90

    def parameters_changed(self):
        self.batch_scale = self._batch_scale()
        ve_active = (not self.stochastic) or self.vem_step
        vm_active = (not self.stochastic) or (not self.vem_step)
        params = flatten_params(self.q_u_means, self.q_u_chols, self.Z, self.kern_list, self.B_list, self.batch_scale,
                                self._W0 if self.compat_stale_W else None, self._k0 if self.compat_stale_W else None)
        self._eng.set_rows(*self._rank_rows())
        out = self._eng.evaluate(params, what="full" if vm_active else "ve")
        self._log_marginal_likelihood = np.array(out["log_marginal"])          # own copies at the reference boundary
        Q = self.num_latent_funcs
        self.q_u_means.gradient = np.array(out["dL_dmu_u"]) if ve_active else np.zeros(self.q_u_means.shape)   # :104-113
        self.q_u_chols.gradient = np.array(out["dL_dL_u"]) if ve_active else np.zeros(self.q_u_chols.shape)
        for q in range(Q):
            if vm_active:                                                                                 # :116-151
                self.kern_list[q].gradient = np.array(out["d_rbf"][q])
                self.B_list[q].gradient = np.concatenate([out["dW"][:, q], out["dkappa"][:, q]])
            else:
                self.kern_list[q].gradient = np.zeros(2)
                self.B_list[q].gradient = np.zeros(2 * self.num_output_funcs)
        if (not self.Z.is_fixed) and vm_active:                                                           # :153-166
            self.Z.gradient = np.array(out["dZ"])
        else:
            self.Z.gradient = np.zeros(self.Z.shape)
        self.status = self._eng.status

    # ------------------------------------------------------------------ paramz-style parameter handling
    def _named(self):
        """(hierarchical name, Param, positive?, variational?) in link order: Z, m_u, L_u, kernels..., B's...
        (svmogp.py:71-75)."""
        out = [(self.name + '.inducing_inputs', self.Z, False, False), (self.name + '.m_u', self.q_u_means, False, True),
               (self.name + '.L_u', self.q_u_chols, False, True)]
        for k in self.kern_list:
            out += [('%s.%s.variance' % (self.name, k.name), k.variance, True, False),
                    ('%s.%s.lengthscale' % (self.name, k.name), k.lengthscale, True, False)]
        for B in self.B_list:
            out += [('%s.%s.W' % (self.name, B.name), B.W, False, False),
                    ('%s.%s.kappa' % (self.name, B.name), B.kappa, True, False)]
        return out

    def _blocks(self):
        return [(p, pos) for _, p, pos, _ in self._named()]

    def __getitem__(self, pattern):
        rx = re.compile(pattern)
        hits = [p for name, p, _, _ in self._named() if rx.match(name)]
        if not hits:
            raise AttributeError("no parameter matches %r" % pattern)
        return _ParamGroup(hits)

    @property
    def optimizer_array(self):
        parts = []
        for p, pos in self._blocks():
            if p.is_fixed:
                continue
            v = np.asarray(p, dtype=np.float64).ravel()
            parts.append(_logexp_finv(v) if pos else v)
        return np.concatenate(parts) if parts else np.zeros(0)

    @optimizer_array.setter
    def optimizer_array(self, x):
        x = np.asarray(x, dtype=np.float64)
        i = 0
        for p, pos in self._blocks():
            if p.is_fixed:
                continue
            n = p.size
            v = x[i:i + n]
            np.asarray(p)[...] = (_logexp_f(v) if pos else v).reshape(p.shape)
            i += n
        self.parameters_changed()

    def _transformed_gradient(self):
        parts = []
        for p, pos in self._blocks():
            if p.is_fixed:
                continue
            g = np.asarray(p.gradient, dtype=np.float64).ravel()
            if pos:
                g = g * _logexp_gradfactor(np.asarray(p, dtype=np.float64).ravel())
            parts.append(g)
        return np.concatenate(parts) if parts else np.zeros(0)

    def _grads(self, x):
        """paramz Model._grads: set the parameters, return -gradient of the objective's transformed params."""
        self.optimizer_array = x
        return -self._transformed_gradient()

    def objective_function(self):
        return -float(self._log_marginal_likelihood[0, 0])

    def _objective_grads(self, x):
        g = self._grads(x)
        return self.objective_function(), g

    def optimize(self, optimizer=None, messages=False, max_iters=1000, **kw):
        """paramz Model.optimize with its default optimiser (L-BFGS-B through scipy, maxfun = maxiter = max_iters), over the
        unfixed parameters.  Every objective / gradient evaluation is one engine call."""
        from scipy import optimize as sopt
        x0 = self.optimizer_array
        if x0.size == 0:
            return None
        res = sopt.fmin_l_bfgs_b(self._objective_grads, x0, iprint=1 if messages else -1, maxfun=max_iters, maxiter=max_iters)
        self.optimizer_array = res[0]
        return res

    def stochastic_grad(self, parameters):                                                               # svmogp.py:188-199
        self.set_data(*self.new_batch())
        stochastic_gradients = self._grads(parameters)
        self._advance_vem()
        return stochastic_gradients

    def _advance_vem(self):
        if self.vem_step:
            if self.ve_count > 2:
                self.ve_count = 0
                self.vem_step = False
            else:
                self.ve_count += 1
        else:
            self.vem_step = True

    def callback(self, i, max_iter, verbose=True, verbose_plot=False):                                   # svmogp.py:201-217
        ll = self.log_likelihood()
        self.elbo[i['n_iter'] - 1, 0] = ll[0][0]
        if verbose and i['n_iter'] % 50 == 0:
            print('svi - iteration ' + str(i['n_iter']) + '/' + str(int(max_iter)))
        if i['n_iter'] > max_iter:
            return True
        return False

    # ------------------------------------------------------------------ device-resident stochastic loop (util.py:320-329)
    def svi_device(self, n_iters, step_rate=0.01, momentum=0.9, decay=0.9, offset=1e-4, trace=True):
        """``n_iters`` iterations of climin.Adadelta(model.optimizer_array, model.stochastic_grad, step_rate, momentum) --
        the loop of util.vem_algorithm(stochastic=True) -- with parameters, gradients and optimiser state resident on
        the device.  Per iteration: next minibatch slice, look-ahead + parameter scatter (one kernel), one engine
        evaluation ('ve' in VE steps, 'full' in VM steps: svmogp.py:104-166 zeroes what the other kind of step would
        have produced), gather + Adadelta update (one kernel).  Returns the ELBO trace (n_iters,) -- the value the
        reference's callback stores, i.e. the ELBO at each look-ahead point -- and leaves the model at the last
        look-ahead point with its ``.gradient`` fields filled, like the reference after ``minimize_until``."""
        from .svi import DeviceSVI
        if self._svi is None or not self._svi.matches(self, step_rate, momentum, decay, offset):
            self._svi = DeviceSVI(self, step_rate, momentum, decay, offset)
        return self._svi.run(n_iters, trace=trace)

    # ------------------------------------------------------------------ prediction (svmogp.py:219-370)
    def _params_now(self):
        return flatten_params(self.q_u_means, self.q_u_chols, self.Z, self.kern_list, self.B_list)

    def _raw_predict(self, Xnew, latent_function_ind=None, full_cov=False, kern=None):
        """svmogp.py:219-250: q(u_q) at Xnew -- mean K_x^T K_uu^-1 m_q and variance k_xx - K_x^T (K_uu^-1 - K_uu^-1 S_q K_uu^-1) K_x
        (diagonal only; the reference's full_cov builds the N x N block).  Evaluated as the output function that mixes
        latent q alone (W = e_q, kappa = 0)."""
        if full_cov:
            raise NotImplementedError("full_cov=True builds an N x N matrix; only the marginal variances are provided")
        q = 0 if latent_function_ind is None else int(latent_function_ind)
        params = dict(self._params_now())
        W = np.zeros_like(params["W"])
        W[0, q] = 1.0
        params["W"], params["kappa"] = W, np.zeros_like(params["kappa"])
        m, v = self._eng.predict_f(params, 0, np.asarray(Xnew, dtype=np.float64))
        return m[:, 0:1].copy(), np.abs(v[:, 0:1])

    def _raw_predict_f(self, Xnew, output_function_ind=None, kern_list=None):
        """q(f_d) at Xnew: (mean (N,1), variance (N,1)).  The reference conditions on the N x N posterior of f_d at the
        training inputs (svmogp.py:263-284, O(N^3)); both routes marginalise the same q(U) and agree where the sparse
        approximation is exact -- this one is O(M^2) per point."""
        d = 0 if output_function_ind is None else int(output_function_ind)
        f_ind = self.Y_metadata['function_index'].flatten()
        d_ind = self.Y_metadata['d_index'].flatten()
        m, v = self._eng.predict_f(self._params_now(), int(f_ind[d]), np.asarray(Xnew, dtype=np.float64))
        k = int(d_ind[d])
        return m[:, k:k + 1].copy(), np.abs(v[:, k:k + 1])                      # np.abs as svmogp.py:284

    predictive_new = _raw_predict_f                                             # svmogp.py:286-312
    _raw_predict_stochastic = _raw_predict_f                                    # svmogp.py:314-338

    def predictive(self, Xpred):                                                # svmogp.py:340-358
        f_index = self.Y_metadata['function_index'].flatten()
        m_F_pred, v_F_pred = [], []
        params = self._params_now()
        for t in range(len(self.likelihood.likelihoods_list)):
            m, v = self._eng.predict_f(params, t, np.asarray(Xpred[t], dtype=np.float64))
            m_F_pred.append(m)
            v_F_pred.append(np.abs(v))
        return self.likelihood.predictive(m_F_pred, v_F_pred, self.Y_metadata)

    def negative_log_predictive(self, Xtest, Ytest, num_samples=1000):          # svmogp.py:360-378
        params = self._params_now()
        mu_F_star, v_F_star = [], []
        for t in range(len(self.likelihood.likelihoods_list)):
            m, v = self._eng.predict_f(params, t, np.asarray(Xtest[t], dtype=np.float64))
            mu_F_star.append(m)
            v_F_star.append(np.abs(v))
        return self.likelihood.negative_log_predictive(Ytest, mu_F_star, v_F_star, Y_metadata=self.Y_metadata,
                                                       num_samples=num_samples)
