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


@pytest.mark.parametrize("header", ["ltephy_b200.h", "ltephy_search.h", "ltephy_sinks.h"])
def test_every_declared_symbol_is_exported(phylib, header):
    names = declared_functions(header)
    assert len(names) >= 7
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


def test_srsran_compat_shim_exports(phylib):
    """tier-2 library: loads on top of libltephy_b200.so and exports every srsran_* function its header declares"""
    from ltesniffer_b200 import build
    assert os.path.exists(build.COMPAT_OUT)
    lib = C.CDLL(build.COMPAT_OUT)
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ltephy_srsran_compat.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(srsran_[a-z0-9_]+)\s*\(", txt)))
    assert names == ["srsran_pdcch_dci_decode", "srsran_ue_dl_decode_fft_estimate", "srsran_ue_dl_decode_pdsch", "srsran_ue_dl_free", "srsran_ue_dl_init",
                     "srsran_ue_dl_set_cell"]
    assert all(hasattr(lib, n) for n in names)
