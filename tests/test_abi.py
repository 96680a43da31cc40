"""CPU-only: the C-ABI library loads, exports every symbol include/agf2_b200.h declares, and fails
loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "agf2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)) - {"defined"})


def test_header_symbols_exported():
    from auxgf_b200.lib import agf2 as lib
    names = _declared_symbols()
    assert "build_part_loop_rhf" in names and "build_part_loop_uhf" in names
    assert "agf2_b200_moments_rhf_dev" in names and "agf2_b200_compress" in names
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), "symbol %s declared in include/agf2_b200.h but not exported" % n
    assert set(lib.SYMBOLS) <= set(names)


def test_drop_in_signatures_match_reference():
    """argtypes of the void symbols are the reference's (auxgf/lib/agf2.py:28-60)."""
    from auxgf_b200.lib import agf2 as lib
    assert len(lib._libagf2.build_part_loop_rhf.argtypes) == 12
    assert len(lib._libagf2.build_part_loop_uhf.argtypes) == 17
    assert lib._libagf2.build_part_loop_rhf.restype is None


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from auxgf_b200.lib import agf2 as lib
    with pytest.raises(lib.Agf2B200Error):
        lib.moments_rhf(np.zeros((4, 4)), np.zeros((4, 4)), np.zeros(2), np.zeros(2), 2)
    with pytest.raises(lib.Agf2B200Error):
        lib.compress(np.eye(3), np.eye(3))
