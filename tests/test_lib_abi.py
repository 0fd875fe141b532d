"""CPU: the C-ABI shared library loads and exports every symbol include/hetmogp_b200.h declares; host-only entry
points (metadata) work without a GPU; compute entry points fail loudly without one (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "hetmogp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hmogp_[a-z_0-9A-Z]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from hetmogp_b200 import _lib
    names = header_symbols()
    assert len(names) >= 20
    raw = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "missing export " + n
    assert set(_lib.EXPORTED) == set(names)
    assert _lib.lib.hmogp_abi_version() == 1


@pytest.mark.parametrize("name", gu.CASES)
def test_c_metadata_bit_exact(name):
    """HetLikelihood.generate_metadata (het_likelihood.py:24-44) through the C-ABI == reference golden."""
    from hetmogp_b200.het_likelihood import HetLikelihood
    from hetmogp_b200 import likelihoods as L
    prob, g = gu.load_case(name)
    het = HetLikelihood([L.from_spec(s) for s in prob["lik_specs"]])
    meta = het.generate_metadata()
    for k in ("task_index", "y_index", "function_index", "d_index", "pred_index"):
        assert meta[k].dtype == g["meta_" + k].dtype or meta[k].dtype.kind == "i"
        assert np.array_equal(np.asarray(meta[k]).ravel(), g["meta_" + k].ravel()), k
    assert het.num_output_functions(meta) == prob["J"]


def test_no_cpu_fallback():
    from hetmogp_b200 import _lib, Engine
    if _lib.lib.hmogp_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.HetMOGPError, match="no CUDA device"):
        Engine([("Bernoulli",)], 8, 1, 1)
    d = _lib.lik_desc(("Bernoulli",))
    y = np.zeros(4)
    rc = _lib.lib.hmogp_lik_var_exp(C.byref(d), 4, y.ctypes.data, y.ctypes.data, y.ctypes.data, None, None, None, 0, 0, None)
    assert rc == _lib.ERR_CUDA and "no CPU fallback" in _lib.last_error()


def test_shard_rows_partition():
    from hetmogp_b200 import shard_rows
    N = [10, 7, 1000003]
    for world in (1, 2, 3, 8):
        covered = [0] * len(N)
        for r in range(world):
            b, c = shard_rows(N, r, world)
            for t in range(len(N)):
                assert b[t] == covered[t]
                covered[t] += c[t]
        assert covered == N
