"""CPU oracle for the HetMOGP ELBO/gradient hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never on the CUDA product path (``hetmogp_b200`` fails
loudly when its CUDA library is missing; it has no CPU fallback).

Parity status: **parity unpinned by reference-shipped tests** -- the reference
(pmorenoz/HetMOGP) ships no tests, fixtures or golden vectors (SURVEY.md §4,
§8c).  The oracle is instead pinned by outputs of the reference itself:
``oracle/verbatim.py`` imports the reference's own hot-path files unmodified
from ``/root/reference`` (over the GPy stand-in in ``oracle/gpy_standin.py``)
and ``oracle/make_golden.py`` stores its results on seeded inputs under
``tests/golden/``.  The travelling restatement (``oracle/diag_oracle.py``,
``oracle/likelihoods_np.py``, ``oracle/params_changed.py``) is checked against
those fixtures in ``tests/test_oracle_golden.py``.

Modules
-------
gpy_standin    minimal GPy/paramz/climin/matplotlib symbols the reference touches
verbatim       loader executing the reference files unmodified (container only)
likelihoods_np numpy restatement of likelihoods/*.py hot-path methods
diag_oracle    diag-only (O(N M^2)) restatement of svmogp_inf.py, row-chunked
params_changed restatement of svmogp.py:85-166 (hyper-parameter chain rule)
synth          seeded synthetic inputs of SURVEY.md §8(d)
make_golden    writes tests/golden/*.npz from the verbatim reference
"""
