"""CPU: libfrt_b200.so loads and exports exactly the functions include/frt.h declares; the ctypes
binding table covers them all; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "frt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from friture_b200 import _lib
    return _lib


def test_every_declared_symbol_is_exported(built):
    names = header_functions()
    assert len(names) >= 15
    lib = ctypes.CDLL(built.lib_path())
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n


def test_binding_table_matches_header(built):
    assert sorted(built.SIGNATURES) == header_functions()


def test_version_and_loud_failure_without_gpu(built):
    lib = built.load_library()
    assert lib.frt_version() >= 100
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.FrtError) as ei:
        built.Handle(0)
    assert "no CPU fallback" in str(ei.value)
    from friture_b200 import audioproc
    import numpy as np
    p = audioproc()
    p.set_fftsize(1024)
    with pytest.raises(built.FrtError):
        p.analyzelive(np.zeros(1024))


def test_no_oracle_import_in_product():
    """Nothing under friture_b200/ (Python or CUDA/C++ sources) names the oracle: it is test
    infrastructure, used only by tests/, __graft_entry__.smoke() and bench.py's checker / CPU legs."""
    pkg = os.path.join(ROOT, "friture_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.lower(), os.path.join(base, f)
