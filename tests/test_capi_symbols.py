"""The C-ABI library loads here (no GPU) and exports every symbol include/odinn_b200.h declares;
without a device the product path fails loudly instead of falling back to the CPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ob():
    import __graft_entry__ as ge

    ge.build_cuda()
    import odinn_b200

    return odinn_b200


def test_every_declared_symbol_is_exported_and_bound(ob):
    hdr = open(os.path.join(ROOT, "include", "odinn_b200.h")).read()
    declared = set(re.findall(r"\b(odinn_[a-z0-9_A-Z]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = ob.load()
    from odinn_b200 import _capi

    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _capi.SIGNATURES, f"{name} has no ctypes prototype"
    assert set(_capi.SIGNATURES) <= declared


def test_no_cpu_fallback(ob, has_cuda):
    if has_cuda:
        pytest.skip("a CUDA device is present")
    with pytest.raises(ob.OdinnError, match="no CUDA device|CUDA"):
        ob.Ensemble([16], [16], [50.0], [50.0])


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "odinn.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                src = open(os.path.join(dp, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|#include\s+[<\"].*oracle", src, re.M), f"{f} uses the oracle"


def test_python_wrappers_refuse_wrong_shapes_before_the_c_call():
    """The C side sees a pointer and a leading dimension only: a transposed / mis-sized matrix or a wrong `out=` must raise
    OdinnError in the wrapper (no handle, hence no GPU, is needed to exercise the checks)."""
    import numpy as np
    import pytest

    import odinn_b200 as ob
    from odinn_b200.ensemble import Ensemble

    ens = Ensemble.__new__(Ensemble)
    ens.G, ens.nx, ens.ny, ens.np_dtype, ens._h = 2, [5, 7], [6, 4], np.float32, None
    assert ens._mat(0, np.zeros((5, 6))).flags.f_contiguous and ens._mat(0, np.zeros((5, 6))).dtype == np.float32
    assert ens._mat(1, np.zeros((6, 3)), dual=True).shape == (6, 3)
    for bad in (np.zeros((6, 5)), np.zeros((5, 5)), np.zeros(30)):
        with pytest.raises(ob.OdinnError):
            ens._mat(0, bad)
    with pytest.raises(ob.OdinnError):
        ens._mat(2, np.zeros((5, 6)))
    good = np.zeros((5, 6), dtype=np.float32, order="F")
    assert ens._out(0, good) is good
    for bad in (np.zeros((5, 6), dtype=np.float64, order="F"), np.zeros((5, 6), dtype=np.float32, order="C"), np.zeros((6, 5), dtype=np.float32, order="F")):
        with pytest.raises(ob.OdinnError):
            ens._out(0, bad)
