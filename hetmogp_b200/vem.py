"""Optimisation drivers of the reference (hetmogp/util.py:284-331) over the CUDA engine.

``vem_algorithm(model, stochastic, vem_iters, step_rate, verbose, optZ, verbose_plot, non_chained)`` keeps the
reference's signature and control flow:

* non-stochastic: ``vem_iters`` rounds of a variational E-step (q(U) free, all hyper-parameters fixed) and a variational
  M-step (kernel variances / lengthscales, W and -- with ``optZ`` -- Z free, q(U) fixed), each one
  ``model.optimize(max_iters=100)`` (util.py:296-318); kappa stays fixed throughout (util.py:289);
* stochastic: climin.Adadelta(model.optimizer_array, model.stochastic_grad, step_rate=0.01, momentum=0.9) run until
  ``model.callback`` stops it, i.e. for ``vem_iters + 1`` gradient evaluations (util.py:320-329, svmogp.py:201-217).
  By default that loop runs resident on the GPU (``SVMOGP.svi_device``: csrc/optim.cu holds the Adadelta state, the
  engine leaves the gradients on the device); ``device_loop=False`` drives the same arithmetic from the host through
  ``hetmogp_b200.optim.Adadelta`` and ``model.stochastic_grad``, callback by callback.
"""
from functools import partial

import numpy as np

_HYPERS = ('.*.lengthscale', '.*.variance', '.*.W')


def _e_step_masks(model):
    for pat in _HYPERS:
        model[pat].fix()
    model.Z.fix()
    model.q_u_means.unfix()
    model.q_u_chols.unfix()


def _m_step_masks(model, optZ, non_chained):
    model['.*.lengthscale'].unfix()
    model['.*.variance'].unfix()
    if optZ:
        model.Z.unfix()
    if non_chained:
        model['.*.W'].unfix()
    model.q_u_means.fix()
    model.q_u_chols.fix()


def vem_algorithm(model, stochastic=False, vem_iters=None, step_rate=None, verbose=False, optZ=True, verbose_plot=False,
                  non_chained=True, device_loop=True):
    model['.*.lengthscale'].fix()                       # util.py:285
    vem_iters = 5 if vem_iters is None else vem_iters
    model['.*.kappa'].fix()                             # util.py:289: "must be always fixed"
    model.elbo = np.empty((vem_iters, 1))
    if stochastic is False:
        for i in range(vem_iters):
            _e_step_masks(model)
            model.optimize(messages=verbose, max_iters=100)
            print('iteration (' + str(i + 1) + ') VE step, ELBO=' + str(model.log_likelihood().flatten()))
            _m_step_masks(model, optZ, non_chained)
            model.optimize(messages=verbose, max_iters=100)
            print('iteration (' + str(i + 1) + ') VM step, ELBO=' + str(model.log_likelihood().flatten()))
        return model
    step_rate = 0.01 if step_rate is None else step_rate
    n_eval = vem_iters + 1                              # the callback stops at n_iter > max_iter (svmogp.py:214-216)
    model.elbo = np.empty((n_eval, 1))
    if device_loop and hasattr(model, "svi_device") and not verbose_plot:
        model.elbo[:, 0] = model.svi_device(n_eval, step_rate=step_rate, momentum=0.9)
        if verbose:
            for it in range(50, n_eval + 1, 50):
                print('svi - iteration ' + str(it) + '/' + str(int(vem_iters)))
    else:
        from .optim import Adadelta
        optimizer = Adadelta(model.optimizer_array, model.stochastic_grad, step_rate=step_rate, momentum=0.9)
        optimizer.minimize_until(partial(model.callback, max_iter=vem_iters, verbose=verbose, verbose_plot=verbose_plot))
    return model
