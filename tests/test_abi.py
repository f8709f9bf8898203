"""CPU tier: the C-ABI library builds, loads and exports every symbol include/dmlmc.h declares;
the product refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dmlmc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmlmc_[a-z0-9_]+)\s*\(", src)))


def test_build_and_symbols():
    import __graft_entry__ as ge
    ge.build()
    from deflatedmlmc_schwinger_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "missing export " + name
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.load().dmlmc_abi_version() == 6


def test_header_cites_reference():
    src = open(os.path.join(ROOT, "include", "dmlmc.h")).read()
    for cite in ("multigrid.py:552", "multigrid.py:406", "multigrid.py:429", "multigrid.py:413", "multigrid.py:347",
                 "utils.py:207", "utils.py:224", "utils.py:213"):
        assert cite in src


def test_no_cpu_fallback():
    import torch
    from deflatedmlmc_schwinger_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.Hierarchy(3)
    # the C ABI itself also refuses
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.dmlmc_hier_create(0, None, 3, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in lib.dmlmc_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deflatedmlmc_schwinger_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
