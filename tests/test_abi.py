"""CPU: the C-ABI library loads, exports every symbol include/ofg.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    from orthofinder_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ofg.h")).read()
    declared = sorted(set(re.findall(r"\b(ofg_[a-z0-9_]+)\s*\(", hdr)))
    assert sorted(_lib.SYMBOLS) == declared
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    import torch
    from orthofinder_b200 import waterfall
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(waterfall.OfgError):
        waterfall.GraphBuilder(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "orthofinder_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/csrc/svml_log10.h", ""), f
