"""hetmogp_b200 -- B200-native engine for the HetMOGP ELBO/gradient hot path.

Drop-in for ONE path of pmorenoz/HetMOGP: ``SVMOGP.parameters_changed()`` -> ``SVMOGPInf.inference`` and the
likelihood ``var_exp`` plug-ins (reference: hetmogp/svmogp.py:85-166, hetmogp/svmogp_inf.py:23-109,
likelihoods/*.py).  All arithmetic runs in hand-written sm_100a CUDA kernels behind the C-ABI of
``include/hetmogp_b200.h``; there is no CPU fallback (importing fails if the shared library is missing).
"""
from . import _lib  # noqa: F401  (fails loudly if libhetmogp_b200.so is not built)
from .engine import Engine, shard_rows  # noqa: F401

__all__ = ["Engine", "shard_rows"]
