"""Line-by-line restatement of SVMOGP.parameters_changed (svmogp.py:85-166).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

hetmogp/svmogp.py cannot be imported without GPy (it subclasses
GPy.core.SparseGP, svmogp.py:16), so its gradient assembly is restated here on
the *dense* ``gradients`` dict the verbatim SVMOGPInf.inference returns
(dL_dKmn as Q x J dense (M, N_t) arrays).  It is the small-N check for the
fused statistics used by oracle/diag_oracle.py and by the CUDA engine.

kern objects are the stand-in RBFs of oracle/gpy_standin.py (GPy semantics
recalled).
"""
import numpy as np

from . import gpy_standin as gpy


def assemble(gradients, problem, Y_metadata, W_chain=None, kappa_chain=None, stochastic=False, vem_step=True,
             Z_fixed=False):
    """Return dict(m_u, L_u, rbf[Q,2], W[J,Q], kappa[J,Q], Z[M,Q*Xdim]) of
    gradients exactly as parameters_changed leaves them in the Param.gradient
    fields."""
    X, Z = problem["X"], problem["Z"]
    Q, Xdim = problem["Q"], problem["Xdim"]
    W, kappa = problem["W"], problem["kappa"]
    Wc = W if W_chain is None else W_chain
    kc = kappa if kappa_chain is None else kappa_chain
    f_index = Y_metadata["function_index"].flatten()
    D = f_index.shape[0]
    M = Z.shape[0]
    kern_list = [gpy.RBF(Xdim, variance=problem["rbf_var"][q], lengthscale=problem["rbf_ls"][q]) for q in range(Q)]
    # self.B_list: the live coregionalisation params; B_list: rebuilt from W_list (svmogp.py:98-99)
    B_live = [gpy.Coregionalize(Xdim, D, 1, W=W[:, q:q + 1], kappa=kappa[:, q]) for q in range(Q)]
    B_loc = [gpy.Coregionalize(Xdim, D, 1, W=Wc[:, q:q + 1], kappa=kc[:, q]) for q in range(Q)]
    out = {"m_u": np.zeros((M, Q)), "L_u": np.zeros((M * (M + 1) // 2, Q)), "rbf": np.zeros((Q, 2)),
           "W": np.zeros((D, Q)), "kappa": np.zeros((D, Q)), "Z": np.zeros_like(Z)}
    Z_grad = np.zeros_like(Z)
    for q, kern_q in enumerate(kern_list):
        Zq = Z[:, q * Xdim:q * Xdim + Xdim]
        ve_active = (not stochastic) or vem_step
        vm_active = (not stochastic) or (not vem_step)
        if ve_active:                                                       # :104-113
            out["m_u"][:, q:q + 1] = gradients["dL_dmu_u"][q]
            out["L_u"][:, q:q + 1] = gradients["dL_dL_u"][q]
        kern_q.update_gradients_full(gradients["dL_dKmm"][q], Zq)          # :116
        grad = kern_q.gradient.copy()
        Kffdiag, KuqF = [], []
        for d in range(D):                                                  # :122-124
            Kffdiag.append(gradients["dL_dKdiag"][q][d])
            KuqF.append(gradients["dL_dKmn"][q][d] * kern_q.K(Zq, X[f_index[d]]))
        # util.update_gradients_diag (util.py:228-231)
        small = np.array([k.sum() for k in Kffdiag])
        Bgrad = np.concatenate([(np.asarray(B_live[q].W) * small[:, None]).ravel(), small])
        # util.update_gradients_Kmn (util.py:248-255)
        dW = np.array([KuqF[d].sum() for d in range(D)])
        Bgrad = Bgrad + np.concatenate([dW, np.zeros(D)])
        if vm_active:                                                       # :130-137
            out["W"][:, q] = Bgrad[:D]
            out["kappa"][:, q] = Bgrad[D:]
        for d in range(D):                                                  # :139-151
            kern_q.update_gradients_full(gradients["dL_dKmn"][q][d], Zq, X[f_index[d]])
            grad += np.asarray(B_loc[q].W)[d] * kern_q.gradient.copy()
            kern_q.update_gradients_diag(gradients["dL_dKdiag"][q][d], X[f_index[d]])
            grad += B_loc[q].B[d, d] * kern_q.gradient.copy()
        if vm_active:
            out["rbf"][q] = grad
        if not Z_fixed:                                                     # :153-156
            Z_grad[:, q * Xdim:q * Xdim + Xdim] += kern_q.gradients_X(gradients["dL_dKmm"][q], Zq)
            for d in range(D):
                Z_grad[:, q * Xdim:q * Xdim + Xdim] += np.asarray(B_loc[q].W)[d] * kern_q.gradients_X(
                    gradients["dL_dKmn"][q][d], Zq, X[f_index[d]])
    if (not Z_fixed) and ((not stochastic) or (not vem_step)):              # :158-166
        out["Z"] = Z_grad
    return out
