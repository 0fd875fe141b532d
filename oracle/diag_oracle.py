"""Diag-only fp64 numpy restatement of the reference ELBO/gradient hot path.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (see oracle/__init__.py).  Travels to
the GPU box (pure numpy/scipy, no reference access).

The literal reference is O(N^2) in time and memory per step because it builds
K_ff and S_fd (N x N) for Posterior objects training never reads
(util.py:176-178, svmogp_inf.py:202,209,219).  Every quantity that reaches the
ELBO and the gradients only uses their diagonals, so this restatement keeps the
reference's formulas and drops the N x N blocks; rows are processed in chunks.
It agrees with the verbatim reference to <=1e-12 on the golden fixtures
(tests/test_oracle_golden.py).

Line map (all into /root/reference/hetmogp/):
  latent_funs_cov         util.py:181-200
  K_fu, W mix             util.py:145-164, svmogp_inf.py:205-218
  diag K_ff               util.py:166-179 (diagonal only), svmogp_inf.py:203,210
  var_exp / derivatives   svmogp_inf.py:73-78, het_likelihood.py:101-131
  KL                      svmogp_inf.py:227-250
  q(U) gradients          svmogp_inf.py:111-183
  hyper-parameter chain   svmogp.py:100-166, util.py:228-255 (see params_changed.py
                          for the line-by-line restatement on dense dL_dKmn)
"""
import numpy as np
from scipy.linalg import lapack

from . import likelihoods_np as lk


def rbf_K(X, X2, variance, lengthscale, same=False):
    """GPy RBF.K (recalled): r^2 via |x|^2+|x'|^2-2x.x', clipped at 0."""
    X1sq = np.sum(np.square(X), 1)
    X2sq = np.sum(np.square(X2), 1)
    r2 = -2.0 * X.dot(X2.T) + (X1sq[:, None] + X2sq[None, :])
    if same:
        r2[np.diag_indices(r2.shape[0])] = 0.0
    r2 = np.clip(r2, 0, np.inf) / lengthscale ** 2
    return variance * np.exp(-0.5 * r2), r2


def jitchol(A, maxtries=5):
    """GPy jitchol semantics (util.py:198).  Returns (L, jitter_used)."""
    L, info = lapack.dpotrf(np.ascontiguousarray(A), lower=1)
    if info == 0:
        return np.tril(L), 0.0
    jitter = np.diag(A).mean() * 1e-6
    for _ in range(maxtries):
        L, info = lapack.dpotrf(np.ascontiguousarray(A + np.eye(A.shape[0]) * jitter), lower=1)
        if info == 0:
            return np.tril(L), jitter
        jitter *= 10
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def chol_inv(L):
    """GPy dpotri: inverse from the lower Cholesky factor, symmetrised."""
    R, _ = lapack.dpotri(np.asfortranarray(L), lower=1)
    return np.tril(R) + np.tril(R, -1).T


def unpack_lower(flat, M):
    """choleskies.flat_to_triang for one column (row-major lower order)."""
    L = np.zeros((M, M))
    L[np.tril_indices(M)] = flat
    return L


def pack_lower(A):
    """choleskies.triang_to_flat for one matrix (reads the lower triangle only)."""
    return A[np.tril_indices(A.shape[0])].copy()


def latent_funs_cov(Z, rbf_var, rbf_ls, Q, Xdim):
    """util.py:181-200."""
    M = Z.shape[0]
    Kuu = np.empty((Q, M, M))
    Luu = np.empty((Q, M, M))
    Kuui = np.empty((Q, M, M))
    jit = np.zeros(Q)
    for q in range(Q):
        Zq = Z[:, q * Xdim:(q + 1) * Xdim]
        Kuu[q], _ = rbf_K(Zq, Zq, rbf_var[q], rbf_ls[q])
        Luu[q], jit[q] = jitchol(Kuu[q])
        Kuui[q] = chol_inv(Luu[q])
    return Kuu, Luu, Kuui, jit


def elbo_and_grads(problem, chunk=8192, W_chain=None, kappa_chain=None, want_rows=False, want_hyper=True,
                   row_slices=None):
    """One evaluation equivalent to SVMOGP.parameters_changed() (svmogp.py:85-166):
    ELBO and all gradients.  ``W_chain``/``kappa_chain`` are the multipliers
    svmogp.py:141,143,156 take from the locally rebuilt B_list (quirk C-5;
    default: the current W, kappa).  ``row_slices`` (list[T] of slices) restricts
    the data term to a row shard -- the per-rank statistic of the multi-GPU
    path; KL terms are still included once."""
    X, Y = problem["X"], problem["Y"]
    Z, m_u, Lflat = problem["Z"], problem["m_u"], problem["L_u"]
    rbf_var, rbf_ls, W, kappa = problem["rbf_var"], problem["rbf_ls"], problem["W"], problem["kappa"]
    Q, Xdim, T = problem["Q"], problem["Xdim"], len(Y)
    M = Z.shape[0]
    Wc = W if W_chain is None else W_chain
    kc = kappa if kappa_chain is None else kappa_chain
    batch_scale = problem.get("batch_scale") or [1.0] * T
    liks = [lk.make(s) for s in problem["lik_specs"]]
    meta = lk.generate_metadata(liks)
    f_index, d_index = meta["function_index"], meta["d_index"]
    J = f_index.shape[0]

    Kuu, Luu, Kuui, jit = latent_funs_cov(Z, rbf_var, rbf_ls, Q, Xdim)
    L_u = np.stack([unpack_lower(Lflat[:, q], M) for q in range(Q)])
    S_u = np.stack([L_u[q].dot(L_u[q].T) for q in range(Q)])
    alpha = np.stack([Kuui[q].dot(m_u[:, q]) for q in range(Q)])          # K_uu^-1 m_q
    SK = np.stack([S_u[q].dot(Kuui[q]) for q in range(Q)])                # S K_uu^-1
    Cq = np.stack([Kuui[q].dot(SK[q]) - Kuui[q] for q in range(Q)])       # K^-1 S K^-1 - K^-1
    kdiag = np.array([sum((W[d, q] ** 2 + kappa[d, q]) * rbf_var[q] for q in range(Q)) for d in range(J)])

    VE_sum = np.zeros(T)
    dVE_dmu = np.zeros((Q, M))
    dVE_dS = np.zeros((Q, M, M))
    sdv = np.zeros(J)                 # sum_n dv_d
    sma = np.zeros((J, Q))            # sum_n dm_d a_tq
    svc = np.zeros((J, Q))            # sum_n dv_d c_tq
    d_ls_mn = np.zeros(Q)
    dZ_mn = np.zeros((M, Q * Xdim))
    n_neg = 0
    rows = {"m": [[] for _ in range(T)], "v": [[] for _ in range(T)], "ve": [[] for _ in range(T)],
            "dm": [[] for _ in range(T)], "dv": [[] for _ in range(T)]}

    for t in range(T):
        ds = np.nonzero(f_index == t)[0]
        sl = slice(0, X[t].shape[0]) if row_slices is None else row_slices[t]
        Xt_all, Yt_all = X[t][sl], Y[t][sl]
        for s in range(0, Xt_all.shape[0], chunk):
            Xc, Yc = Xt_all[s:s + chunk], Yt_all[s:s + chunk]
            n = Xc.shape[0]
            Ks, r2s, As, a, c = [], [], [], np.empty((Q, n)), np.empty((Q, n))
            for q in range(Q):
                Zq = Z[:, q * Xdim:(q + 1) * Xdim]
                K, r2 = rbf_K(Xc, Zq, rbf_var[q], rbf_ls[q])
                A = K.dot(Kuui[q])                                        # svmogp_inf.py:214-215 (unscaled by W)
                a[q] = A.dot(m_u[:, q])                                   # :216
                c[q] = np.sum(np.square(A.dot(L_u[q])), 1) - np.sum(A * K, 1)  # :217-218
                Ks.append(K), r2s.append(r2), As.append(A)
            Mf = np.stack([sum(W[d, q] * a[q] for q in range(Q)) for d in ds], axis=1)
            Vf = np.stack([kdiag[d] + sum(W[d, q] ** 2 * c[q] for q in range(Q)) for d in ds], axis=1)
            n_neg += int((Vf < 0).sum())
            ve = liks[t].var_exp(Yc, Mf, Vf) * batch_scale[t]             # :73,76
            dm, dv = liks[t].var_exp_derivatives(Yc, Mf, Vf)              # :74
            dm, dv = dm * batch_scale[t], dv * batch_scale[t]             # :77-78
            VE_sum[t] += ve.sum()
            if want_rows:
                for k, arr in (("m", Mf), ("v", Vf), ("ve", ve), ("dm", dm), ("dv", dv)):
                    rows[k][t].append(arr)
            sdv[ds] += dv.sum(0)
            for q in range(Q):
                mu_q = dm.dot(W[ds, q])                                   # sum_d W_dq dm_d
                om_q = dv.dot(W[ds, q] ** 2)                              # sum_d W_dq^2 dv_d
                dVE_dmu[q] += As[q].T.dot(mu_q)                           # :144
                dVE_dS[q] += (As[q].T * om_q).dot(As[q])                  # :145-148
                sma[ds, q] += dm.T.dot(a[q])
                svc[ds, q] += dv.T.dot(c[q])
                if want_hyper:
                    # Gamma'[n,m] = sum_d W'_dq dL_dKmn_d[m,n]  (svmogp_inf.py:157-161, svmogp.py:140-141,156)
                    mu_c = dm.dot(Wc[ds, q])
                    om_c = dv.dot(Wc[ds, q] * W[ds, q])
                    CK = Ks[q].dot(Cq[q])                                 # rows (C k_n)^T
                    G = mu_c[:, None] * alpha[q][None, :] + 2.0 * om_c[:, None] * CK
                    GK = G * Ks[q]
                    d_ls_mn[q] += np.sum(GK * r2s[q]) / rbf_ls[q]         # RBF.update_gradients_full, lengthscale
                    Zq = Z[:, q * Xdim:(q + 1) * Xdim]
                    for i in range(Xdim):                                 # RBF.gradients_X(dL_dKmn, Z, X)
                        dZ_mn[:, q * Xdim + i] += (GK * (Xc[:, i][:, None] - Zq[:, i][None, :])).sum(0) / rbf_ls[q] ** 2

    # KL (svmogp_inf.py:243-250)
    KL = 0.0
    for q in range(Q):
        KL += 0.5 * np.sum(Kuui[q] * S_u[q]) + 0.5 * m_u[:, q].dot(Kuui[q]).dot(m_u[:, q]) - 0.5 * M \
            + np.sum(np.log(np.abs(np.diag(Luu[q])))) - np.sum(np.log(np.abs(np.diag(L_u[q]))))
    log_marginal = VE_sum.sum() - KL

    out = {"log_marginal": np.array([[log_marginal]]), "VE_sum": VE_sum, "KL": KL, "n_negative_v": n_neg,
           "jitter": jit, "dL_dmu_u": [], "dL_dL_u": [], "dL_dKmm": [], "Kuu": Kuu, "Luu": Luu, "Kuui": Kuui}
    d_rbf = np.zeros((Q, 2))
    dW = np.zeros((J, Q))
    dkappa = np.zeros((J, Q))
    dZ = dZ_mn.copy()
    for q in range(Q):
        S_qi = chol_inv(L_u[q])                                           # :124
        if np.any(np.isinf(S_qi)):
            raise ValueError("Sqi: Cholesky representation unstable")    # :126-127
        dKL_dmu = alpha[q]
        dKL_dS = 0.5 * (Kuui[q] - S_qi)
        dKL_dK = 0.5 * Kuui[q] - 0.5 * Kuui[q].dot(S_u[q]).dot(Kuui[q]) - 0.5 * np.outer(alpha[q], alpha[q])
        E = dVE_dS[q]
        tmp = E.dot(S_u[q]).dot(Kuui[q])                                  # :151
        dVE_dK = E - tmp - tmp.T - np.outer(dVE_dmu[q], alpha[q])         # :152-154
        dVE_dK = 0.5 * (dVE_dK + dVE_dK.T)                                # :166
        dL_dmu = dVE_dmu[q] - dKL_dmu
        dL_dS = E - dKL_dS
        dL_dK = dVE_dK - dKL_dK
        out["dL_dmu_u"].append(dL_dmu[:, None])
        out["dL_dL_u"].append(pack_lower(2.0 * dL_dS.dot(L_u[q]))[:, None])  # :175-178
        out["dL_dKmm"].append(dL_dK)
        if want_hyper:
            Zq = Z[:, q * Xdim:(q + 1) * Xdim]
            _, r2 = rbf_K(Zq, Zq, rbf_var[q], rbf_ls[q])
            KG = Kuu[q] * dL_dK
            d_rbf[q, 0] = KG.sum() / rbf_var[q]                           # svmogp.py:116
            d_rbf[q, 1] = (KG * r2).sum() / rbf_ls[q]
            # K_mn and K_diag chains (svmogp.py:139-143)
            d_rbf[q, 0] += sum(Wc[d, q] * (sma[d, q] + 2.0 * W[d, q] * svc[d, q]) for d in range(J)) / rbf_var[q]
            d_rbf[q, 1] += d_ls_mn[q]
            d_rbf[q, 0] += sum((Wc[d, q] ** 2 + kc[d, q]) * sdv[d] for d in range(J))
            # W, kappa (util.py:228-231,248-254; svmogp.py:120-129)
            dW[:, q] = W[:, q] * sdv + sma[:, q] + 2.0 * W[:, q] * svc[:, q]
            dkappa[:, q] = sdv
            # Z (svmogp.py:153-156): gradients_X(dL_dKmm, Z) with X2=None
            tmpz = -KG
            tmpz = tmpz + tmpz.T
            for i in range(Xdim):
                dZ[:, q * Xdim + i] += (tmpz * (Zq[:, i][:, None] - Zq[:, i][None, :])).sum(1) / rbf_ls[q] ** 2
    if want_hyper:
        out.update(d_rbf=d_rbf, dW=dW, dkappa=dkappa, dZ=dZ)
    out.update(sdv=sdv, sma=sma, svc=svc, dVE_dmu=dVE_dmu, dVE_dS=dVE_dS)
    if want_rows:
        out["rows"] = {k: [np.concatenate(v) if v else None for v in rows[k]] for k in rows}
    return out
