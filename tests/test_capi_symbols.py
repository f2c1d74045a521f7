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
