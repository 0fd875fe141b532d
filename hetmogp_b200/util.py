"""Model-construction and minibatch helpers with the reference's names (hetmogp/util.py:15-143), host logic only.

The kernel objects are the minimal containers of ``gpy_shim`` (GPy is not importable here); no matrix of the
path is built on the host -- ``cross_covariance`` / ``latent_funs_cov`` of the reference (util.py:145-200) live in
the CUDA kernels (hetmogp_b200/csrc/proj_*.cu, mm_algebra.cu).
"""
import random

import numpy as np

from .gpy_shim import RBF, Coregionalize
from .toy import true_u_functions, true_f_functions, generate_toy_U   # noqa: F401  (util.py:21-50,202-206)
from .vem import vem_algorithm                                          # noqa: F401  (util.py:284-331)


def get_batch_scales(X_all, X):                                             # util.py:15-19
    return [float(xa.shape[0]) / float(X[t].shape[0]) for t, xa in enumerate(X_all)]


def mini_slices(n_samples, batch_size):                                     # util.py:52-60 (bit-exact slice bounds)
    """The last slice is NOT clamped to n_samples (util.py:60): indexing clamps it, the bounds themselves do not."""
    n_batches, rest = divmod(n_samples, batch_size)
    if rest != 0:
        n_batches += 1
    return [slice(i * batch_size, (i + 1) * batch_size) for i in range(n_batches)]


def draw_mini_slices(n_samples, batch_size, with_replacement=False):        # util.py:62-72
    """Same stream as the reference: the shuffle there acts on a temporary copy, so slices are served in order
    0,1,2,... cyclically (quirk C-7)."""
    slices = mini_slices(n_samples, batch_size)
    idxs = list(range(len(slices)))
    if with_replacement:
        yield random.choice(slices)
    else:
        while True:
            random.shuffle(list(idxs))
            for i in idxs:
                yield slices[i]


def latent_functions_prior(Q, lenghtscale=None, variance=None, input_dim=None):   # util.py:75-90
    lenghtscale = np.random.rand(Q) if lenghtscale is None else lenghtscale
    variance = np.random.rand(Q) if variance is None else variance
    kern_list = []
    for q in range(Q):
        kern_q = RBF(input_dim=input_dim, lengthscale=lenghtscale[q], variance=variance[q], name='rbf')
        kern_q.name = 'kern_q' + str(q)
        kern_list.append(kern_q)
    return kern_list


def random_W_kappas(Q, D, rank, experiment=False):                          # util.py:92-104
    W_list, kappa_list = [], []
    for q in range(Q):
        p = np.random.binomial(n=1, p=0.5 * np.ones((D, 1)))
        Ws = p * np.random.normal(loc=0.5, scale=0.5, size=(D, 1)) - (p - 1) * np.random.normal(loc=-0.5, scale=0.5, size=(D, 1))
        W_list.append(Ws / np.sqrt(rank))
        kappa_list.append(np.zeros(D))
    return W_list, kappa_list


def ICM(input_dim, output_dim, kernel, rank, W=None, kappa=None, name='ICM'):   # util.py:106-124
    B = Coregionalize(input_dim=input_dim, output_dim=output_dim, rank=rank, W=W, kappa=kappa)
    B.name = name
    return (kernel, B), B


def LCM(input_dim, output_dim, kernels_list, W_list, kappa_list, rank, name='B_q'):   # util.py:126-143
    K, B_q = [], []
    for q, kernel in enumerate(kernels_list):
        Kq, Bq = ICM(input_dim, output_dim, kernel, W=W_list[q], kappa=kappa_list[q], rank=rank, name='%s%s' % (name, q))
        K.append(Kq)
        B_q.append(Bq)
    return K, B_q

