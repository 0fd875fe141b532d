"""Minimal parameter/kernel containers used when GPy is not importable (it is not in this image).

Only the attributes the hot path reads or writes exist: ``Param`` (ndarray with ``.gradient``/``.values``/fix
flags), ``RBF(input_dim, variance, lengthscale)`` with ``.variance/.lengthscale/.gradient``, and
``Coregionalize(input_dim, output_dim, rank, W, kappa)`` with ``.W/.kappa/.B`` (reference: hetmogp/util.py:75-143,
GPy semantics in SURVEY.md App. D).  No arithmetic of the path lives here -- kernels matrices are built on the GPU.
"""
import numpy as np


class Param(np.ndarray):
    def __new__(cls, name, value):
        obj = np.array(value, dtype=np.float64, copy=True).view(cls)
        obj.name = name
        obj.gradient = np.zeros(obj.shape)
        obj.is_fixed = False
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.name = getattr(obj, "name", None)
        self.is_fixed = getattr(obj, "is_fixed", False)
        if not hasattr(self, "gradient"):
            self.gradient = None

    @property
    def values(self):
        return np.asarray(self)

    def fix(self):
        self.is_fixed = True

    def unfix(self):
        self.is_fixed = False


class RBF(object):
    def __init__(self, input_dim, variance=1.0, lengthscale=1.0, ARD=False, name='rbf'):
        self.input_dim = input_dim
        self.name = name
        self.variance = Param('variance', np.atleast_1d(variance))
        self.lengthscale = Param('lengthscale', np.atleast_1d(lengthscale))
        self._gradient = np.zeros(2)

    @property
    def gradient(self):
        return self._gradient

    @gradient.setter
    def gradient(self, g):
        self._gradient = np.asarray(g, dtype=np.float64).reshape(2)
        self.variance.gradient[...] = self._gradient[0]
        self.lengthscale.gradient[...] = self._gradient[1]

    def copy(self):
        return RBF(self.input_dim, float(self.variance[0]), float(self.lengthscale[0]), name=self.name)


class Coregionalize(object):
    def __init__(self, input_dim, output_dim, rank=1, W=None, kappa=None, name='B'):
        self.input_dim, self.output_dim, self.rank, self.name = input_dim, output_dim, rank, name
        self.W = Param('W', np.zeros((output_dim, rank)) if W is None else np.asarray(W).reshape(output_dim, rank))
        self.kappa = Param('kappa', np.zeros(output_dim) if kappa is None else np.asarray(kappa).reshape(output_dim))
        self._gradient = np.zeros(output_dim * (rank + 1))

    @property
    def B(self):
        return np.asarray(self.W).dot(np.asarray(self.W).T) + np.diag(np.asarray(self.kappa))

    @property
    def gradient(self):
        return self._gradient

    @gradient.setter
    def gradient(self, g):
        self._gradient = np.asarray(g, dtype=np.float64).ravel()
        D = self.output_dim
        self.W.gradient[...] = self._gradient[:D * self.rank].reshape(D, self.rank)
        self.kappa.gradient[...] = self._gradient[D * self.rank:]
