"""SVMOGP model container with the reference's surface (hetmogp/svmogp.py:16-217), backed by the CUDA engine.

Keeps ``SVMOGP(X, Y, Z, kern_list, likelihood, Y_metadata, name, batch_size, W_list)``, ``parameters_changed()``
(writes the ``.gradient`` fields exactly as svmogp.py:100-166 leaves them, including the VE/VM gating of the
stochastic mode), ``log_likelihood()`` (a (1,1) array, svmogp.py:82-83), ``new_batch`` / ``set_data`` /
``stochastic_grad`` / ``callback`` (svmogp.py:168-217) and a paramz-style flat ``optimizer_array`` with the
Logexp transform of positive parameters.  The data stay resident on the GPU; a minibatch is a row slice.
"""
import numpy as np

from .gpy_shim import Param, RBF, Coregionalize
from .svmogp_inf import SVMOGPInf, flatten_params
from . import util


def _logexp_f(x):      # paramz Logexp: theta = log(1 + e^x)
    return np.where(x > 30.0, x, np.log1p(np.exp(np.minimum(x, 30.0))))


def _logexp_finv(t):
    return np.where(t > 30.0, t, np.log(np.expm1(np.minimum(t, 30.0))))


class SVMOGP(object):
    def __init__(self, X, Y, Z, kern_list, likelihood, Y_metadata, name='SVMOGP', batch_size=None, W_list=None,
                 precision="fp32", device=0, group=None, compat_stale_W=False):
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
        # engine with the full data resident; minibatches are row slices (util.py:52-72)
        self._eng = self.inference_method.engine_for(likelihood, M, self.num_latent_funcs, self.Xdim)
        self._eng.set_data(X, Y)
        self._N_all = [x.shape[0] for x in X]
        if batch_size is None:
            self.stochastic = False
            self._slice = [(0, n) for n in self._N_all]
        else:
            self.stochastic = True
            self.slicer_list = [util.draw_mini_slices(n, self.batch_size) for n in self._N_all]
            self.new_batch()
        self.vem_step = True
        self.ve_count = 0
        self.elbo = np.zeros((1, 1))
        self._log_marginal_likelihood = np.zeros((1, 1))
        self.parameters_changed()

    # ------------------------------------------------------------------ data / minibatching (svmogp.py:168-186)
    @property
    def Xmulti(self):
        return [x[b:b + c] for x, (b, c) in zip(self.Xmulti_all, self._slice)]

    @property
    def Ymulti(self):
        return [y[b:b + c] for y, (b, c) in zip(self.Ymulti_all, self._slice)]

    def new_batch(self):
        sl = [next(s) for s in self.slicer_list]
        self._slice = [(s.start, s.stop - s.start) for s in sl]
        return self.Xmulti, self.Ymulti

    def set_data(self, X=None, Y=None):
        pass   # the batch is a slice of the resident data; kept for interface parity (svmogp.py:168-173)

    # ------------------------------------------------------------------ the hot path (svmogp.py:82-166)
    def log_likelihood(self):
        return self._log_marginal_likelihood

    def parameters_changed(self):
        T = len(self.likelihood.likelihoods_list)
        self.batch_scale = [float(self._N_all[t] / self._slice[t][1]) for t in range(T)]                 # svmogp.py:89-90
        ve_active = (not self.stochastic) or self.vem_step
        vm_active = (not self.stochastic) or (not self.vem_step)
        params = flatten_params(self.q_u_means, self.q_u_chols, self.Z, self.kern_list, self.B_list, self.batch_scale,
                                self._W0 if self.compat_stale_W else None, self._k0 if self.compat_stale_W else None)
        self._eng.set_rows([b for b, _ in self._slice], [c for _, c in self._slice])
        out = self._eng.evaluate(params, what="full" if vm_active else "ve")
        self._log_marginal_likelihood = out["log_marginal"]
        Q = self.num_latent_funcs
        self.q_u_means.gradient = out["dL_dmu_u"].copy() if ve_active else np.zeros_like(out["dL_dmu_u"])  # :104-113
        self.q_u_chols.gradient = out["dL_dL_u"].copy() if ve_active else np.zeros_like(out["dL_dL_u"])
        for q in range(Q):
            if vm_active:                                                                                 # :116-151
                self.kern_list[q].gradient = out["d_rbf"][q]
                self.B_list[q].gradient = np.concatenate([out["dW"][:, q], out["dkappa"][:, q]])
            else:
                self.kern_list[q].gradient = np.zeros(2)
                self.B_list[q].gradient = np.zeros(2 * self.num_output_funcs)
        if (not self.Z.is_fixed) and vm_active:                                                           # :153-166
            self.Z.gradient = out["dZ"].copy()
        else:
            self.Z.gradient = np.zeros(self.Z.shape)
        self.status = self._eng.status

    # ------------------------------------------------------------------ paramz-style flat parameter vector
    def _blocks(self):
        """(param, positive?) in link order: Z, m_u, L_u, kernels..., B's... (svmogp.py:71-75)."""
        blocks = [(self.Z, False), (self.q_u_means, False), (self.q_u_chols, False)]
        for k in self.kern_list:
            blocks += [(k.variance, True), (k.lengthscale, True)]
        for B in self.B_list:
            blocks += [(B.W, False), (B.kappa, True)]
        return blocks

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
            if pos:   # Logexp gradient factor 1 - e^{-theta}
                g = g * (1.0 - np.exp(-np.asarray(p, dtype=np.float64).ravel()))
            parts.append(g)
        return np.concatenate(parts) if parts else np.zeros(0)

    def _grads(self, x):
        """paramz Model._grads: set the parameters, return -gradient of the objective's transformed params."""
        self.optimizer_array = x
        return -self._transformed_gradient()

    def objective_function(self):
        return -float(self._log_marginal_likelihood[0, 0])

    def stochastic_grad(self, parameters):                                                               # svmogp.py:188-199
        self.set_data(*self.new_batch())
        stochastic_gradients = self._grads(parameters)
        if self.vem_step:
            if self.ve_count > 2:
                self.ve_count = 0
                self.vem_step = False
            else:
                self.ve_count += 1
        else:
            self.vem_step = True
        return stochastic_gradients

    def callback(self, i, max_iter, verbose=True, verbose_plot=False):                                   # svmogp.py:201-217
        ll = self.log_likelihood()
        self.elbo[i['n_iter'] - 1, 0] = self.log_likelihood()[0]
        if verbose and i['n_iter'] % 50 == 0:
            print('svi - iteration ' + str(i['n_iter']) + '/' + str(int(max_iter)))
        if i['n_iter'] > max_iter:
            return True
        return False
