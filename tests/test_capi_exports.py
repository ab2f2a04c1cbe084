"""The C-ABI library loads (no GPU needed) and exports every function include/*.h declares; creation fails
loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import pytest
from ltesniffer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ltephy_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["ltephy_b200.h", "ltephy_search.h"])
def test_every_declared_symbol_is_exported(phylib, header):
    names = declared_functions(header)
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(phylib, n)]
    assert not missing, "declared in include/%s but not exported: %s" % (header, missing)


def test_create_fails_loudly_without_gpu(phylib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        capi.LtePhy(100, 1, 1, 1)


def test_invalid_inputs_are_rejected(phylib):
    h = C.c_void_p()
    cfg = capi.Cfg(nof_prb=6, nof_ports=1, cell_id=0, nof_rx=1, max_subframes=1)
    assert phylib.ltephy_create(C.byref(cfg), C.byref(h)) == -2   # LTEPHY_ERROR_INVALID_INPUTS (falcon_pdcch.c:121 convention)
    assert phylib.ltephy_create(None, C.byref(h)) == -2
