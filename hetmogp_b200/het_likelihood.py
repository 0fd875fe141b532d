"""HetLikelihood with the reference's interface (hetmogp/het_likelihood.py:10-164)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check


class HetLikelihood(object):
    def __init__(self, likelihoods_list, gp_link=None, name='heterogeneous_likelihood'):
        self.likelihoods_list = likelihoods_list
        self.name = name

    def specs(self):
        return [tuple(l.spec) for l in self.likelihoods_list]

    def generate_metadata(self):
        """het_likelihood.py:24-44 -- integer index maps, computed by the C-ABI (bit-exact)."""
        T = len(self.likelihoods_list)
        descs = (_lib.LikDesc * T)(*[_lib.lik_desc(s) for s in self.specs()])
        ny, nf, npred = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.hmogp_generate_metadata(T, descs, None, None, None, None, None, C.byref(ny), C.byref(nf), C.byref(npred)))
        t_index = np.empty(T, dtype=np.int64)
        y_index = np.empty(ny.value, dtype=np.int64)
        f_index = np.empty(nf.value, dtype=np.int64)
        d_index = np.empty(nf.value, dtype=np.int64)
        p_index = np.empty(npred.value, dtype=np.int64)
        as_p = lambda a: a.ctypes.data_as(_lib.c_int64_p)
        check(lib.hmogp_generate_metadata(T, descs, as_p(t_index), as_p(y_index), as_p(f_index), as_p(d_index), as_p(p_index),
                                          None, None, None))
        return {'task_index': t_index, 'y_index': np.int_(y_index), 'function_index': np.int_(f_index),
                'd_index': np.int_(d_index), 'pred_index': np.int_(p_index)}

    def num_output_functions(self, Y_metadata):
        return Y_metadata['function_index'].flatten().shape[0]   # het_likelihood.py:85-90

    def ismulti(self, task):
        return self.likelihoods_list[task].ismulti()

    def var_exp(self, Y, mu_F, v_F, Y_metadata):
        tasks = np.unique(Y_metadata['task_index'].flatten())
        return [self.likelihoods_list[t].var_exp(Y[t], mu_F[t], v_F[t], Y_metadata=None) for t in tasks]

    def var_exp_derivatives(self, Y, mu_F, v_F, Y_metadata):
        tasks = np.unique(Y_metadata['task_index'].flatten())
        dm, dv = [], []
        for t in tasks:
            a, b = self.likelihoods_list[t].var_exp_derivatives(Y[t], mu_F[t], v_F[t], Y_metadata=None)
            dm.append(a)
            dv.append(b)
        return dm, dv

    def predictive(self, mu_F_pred, v_F_pred, Y_metadata):
        """het_likelihood.py:133-148: per-task predictive mean and variance."""
        tasks = np.unique(Y_metadata['task_index'].flatten())
        m_pred, v_pred = [], []
        for t in tasks:
            m, v = self.likelihoods_list[t].predictive(mu_F_pred[t], v_F_pred[t], Y_metadata=None)
            m_pred.append(m)
            v_pred.append(v)
        return m_pred, v_pred

    def negative_log_predictive(self, Ytest, mu_F_star, v_F_star, Y_metadata, num_samples):
        """het_likelihood.py:150-164: NLPD over the test data of every task."""
        tasks = np.unique(Y_metadata['task_index'].flatten())
        logpred = 0
        for t in tasks:
            logpred += self.likelihoods_list[t].log_predictive(Ytest[t], mu_F_star[t], v_F_star[t], num_samples)
        return -logpred
