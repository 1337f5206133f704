"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol
include/boomgpu.h declares; no compute calls (those need a GPU)."""
import os
import re

import pytest

import boom_b200
from boom_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, "include", "boomgpu.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"\b(boomgpu_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    if not os.path.exists(capi.library_path()):
        pytest.skip("libboomgpu.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    lib = capi.load_library()
    declared = _declared()
    assert len(declared) >= 26
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(capi.SYMBOLS) == declared
    assert b"sm_100a" in lib.boomgpu_version()


def test_no_cpu_fallback():
    """Without a GPU the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    if not os.path.exists(capi.library_path()):
        pytest.skip("libboomgpu.so not built")
    with pytest.raises(boom_b200.BoomGpuError, match="no CPU fallback"):
        boom_b200.Context(0)


def test_product_does_not_import_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "boom_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "auxmix_oracle" not in text and "from oracle" not in text and "import oracle" not in text, fn
