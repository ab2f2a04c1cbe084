"""phich-Resource (Ng) of the MIB: the number of PHICH groups in symbol 0 changes which REGs the PDCCH gets, hence the CCE grid.  The reference's file mode
presets Ng = 1/6 (src/src/LTESniffer_Core.cc:242-247) but its live mode hands srsRAN the MIB's value (srsran_cell_t.phich_resources, :196, :389), so a
drop-in has to take it: ltephy_cfg_t.phich_resources, ltephy_search_create_cell_ng, srsran_ue_dl_set_cell of the compat layer.  The control-region map is
built on the host and uploaded as a table (the kernels only gather through it), so it is checked here without a GPU:
  * group and CCE counts against the closed forms of 36.211 6.9 / 6.8.1,
  * the product's map (ltephy_ctrl_region_map) against the oracle's (an independent C restatement in sim/lte_common.c), RE by RE, for every Ng / CFI,
  * the synthetic eNB -> oracle receiver loop at Ng = 1/2, 1 and 2 (every DCI found again),
  * Ng = 1/6 is the zero value of the field: nothing changes for existing callers,
  * the same for phich-Duration extended (ltephy_cfg_t.phich_length; REGs of every group in symbols 0, 1 and 2)."""
import ctypes as C
import math
import numpy as np
import pytest
import ltelib
from ltelib import Cell, Sim, Oracle
from ltesniffer_b200 import capi

NG = {0: 1 / 6, 1: 1 / 2, 2: 1.0, 3: 2.0}


def product_map(cellp, ng, cfi, ext=0):
    L = capi.load_library()
    capi._bind_search(L)
    idx = np.zeros(88 * 36, np.uint16)
    pc = np.zeros(16, np.uint16)
    n = C.c_uint32(0)
    r = L.ltephy_ctrl_region_map(cellp[0], cellp[1], cellp[2], ng | (ext << 8), cfi, idx.ctypes.data_as(C.c_void_p), len(idx), C.byref(n), pc.ctypes.data_as(C.c_void_p))
    assert r == 0, (cellp, ng, cfi, r)
    return n.value, idx[:36 * n.value].copy(), pc


@pytest.mark.parametrize("cellp", [(100, 2, 301, 2), (75, 2, 17, 2), (50, 1, 9, 1), (25, 2, 150, 2), (15, 1, 503, 1)])
def test_cce_counts_and_map_equal_oracle_for_every_ng(infra, cellp):
    O = infra.oracle()
    O.lteo_pdcch_re_index.restype = C.c_uint32
    O.lteo_pdcch_re_index.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    nprb = cellp[0]
    seen = set()
    for ng in range(4):
        cell = Cell(*cellp, 0, ng)
        o = Oracle(cell)
        groups = math.ceil(NG[ng] * nprb / 8 - 1e-9)                        # 36.211 6.9, normal cyclic prefix
        for cfi in (1, 2, 3):
            nregs = nprb * (2 + (3 if cfi >= 2 else 0) + (3 if cfi >= 3 else 0))    # 1 / 2 antenna ports: 2, 3, 3 REGs per PRB in symbols 0, 1, 2
            want_cce = (nregs - 4 - 3 * groups) // 9                        # PCFICH takes 4 REGs, every PHICH group 3
            n, idx, pc = product_map(cellp, ng, cfi)
            assert n == want_cce == O.lteo_nof_cce(o.h, cfi), (cellp, ng, cfi, n, want_cce)
            oidx = np.zeros(88 * 36, np.uint16)
            opc = np.zeros(16, np.uint16)
            assert O.lteo_pdcch_re_index(o.h, cfi, ltelib.ptr(oidx), ltelib.ptr(opc)) == n
            assert np.array_equal(idx, oidx[:36 * n]) and np.array_equal(pc, opc), (cellp, ng, cfi)
            assert len(set(idx.tolist())) == len(idx) and not set(idx.tolist()) & set(pc.tolist())     # every RE once, none shared with the PCFICH
            assert int(idx.max()) < cfi * 12 * nprb
            seen.add((ng, cfi, n))
    assert len({n for ng, cfi, n in seen if cfi == 3}) >= 2                  # the CCE grid really depends on Ng


def test_default_is_one_sixth(infra):
    L = capi.load_library()
    capi._bind_search(L)
    assert capi.Cfg().phich_resources == 0
    for cfi, want in ((1, 20), (2, 54), (3, 87)):                            # SURVEY.md App. C, 100 PRB, Ng = 1/6
        assert product_map((100, 2, 1, 2), 0, cfi)[0] == want
    assert [product_map((100, 2, 1, 2), 2, cfi)[0] for cfi in (1, 2, 3)] == [17, 50, 84]      # Ng = 1: 13 groups
    a = capi.Search(100, 2, 1, 2)
    b = capi.Search(100, 2, 1, 2, phich_resources=0)
    c = capi.Search(100, 2, 1, 2, phich_resources=3)
    assert a.h and b.h and c.h
    idx = np.zeros(4, np.uint16)
    n = C.c_uint32(0)
    assert L.ltephy_ctrl_region_map(100, 2, 1, 4, 1, None, 0, C.byref(n), None) == -2       # no such phich-Resource
    assert L.ltephy_search_create_cell_ng(100, 2, 1, 2, 4, 5) is None


@pytest.mark.parametrize("ng", [1, 2, 3])
def test_transmitter_to_oracle_loop_with_other_phich_resources(infra, ng):
    """the synthetic eNB maps its PDCCH around the PHICH groups of this Ng; the oracle receiver, told the same Ng, finds every DCI again -- and a receiver
    that assumes 1/6 does not (the CCE grid differs)"""
    from helpers import make_capture
    cell = Cell(50, 2, 21, 2, 0, ng)
    sim, iq, tti, truths, payloads = make_capture(cell, 6, seed=30 + ng, cfi=2, nof_ues=6, dl_min=3, dl_max=5, tm=2, mcs_min=4, mcs_max=12, snr_db=25.0)
    o = Oracle(cell)
    o16 = Oracle(Cell(50, 2, 21, 2, 0, 0))
    found = total = found16 = 0
    for sf in range(6):
        sym = o.ofdm(iq[sf])
        ce, res = o.chest(int(tti[sf]) % 10, sym)
        cfi, corr = o.pcfich(int(tti[sf]) % 10, sym, ce)
        assert cfi == truths[sf].cfi
        llr = o.pdcch_llr(int(tti[sf]) % 10, cfi, sym, ce)
        ce16, _ = o16.chest(int(tti[sf]) % 10, sym)
        llr16 = o16.pdcch_llr(int(tti[sf]) % 10, cfi, sym, ce16)
        for i in range(truths[sf].nof_dci):
            d = truths[sf].dci[i]
            total += 1
            for oo, ll, which in ((o, llr, 0), (o16, llr16, 1)):
                if 72 * (d.ncce + (1 << d.L)) > len(ll):
                    continue
                r, bits, crc = oo.dci_decode(ll[72 * d.ncce:72 * (d.ncce + (1 << d.L))], d.nbits)
                ok = r == 0 and crc == d.rnti and np.array_equal(bits, np.frombuffer(bytes(d.bits), np.uint8)[:d.nbits])
                if which == 0:
                    found += ok
                else:
                    found16 += ok
    assert total >= 15 and found == total and found16 < total // 2, (total, found, found16)


@pytest.mark.parametrize("cellp", [(100, 2, 301, 2), (75, 2, 17, 2), (50, 1, 9, 1), (25, 2, 150, 2), (15, 1, 503, 1)])
def test_extended_phich_duration_map_equals_oracle(infra, cellp):
    """phich-Duration extended: one REG of every PHICH group in each of the first three symbols (36.211 6.9.3), so symbol 0 loses one REG per group instead of
    three and symbols 1 and 2 one each: CCE counts by closed form, the product's map against the oracle's RE by RE for every Ng and CFI"""
    O = infra.oracle()
    O.lteo_pdcch_re_index.restype = C.c_uint32
    O.lteo_pdcch_re_index.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    nprb = cellp[0]
    for ng in range(4):
        o = Oracle(Cell(*cellp, 0, ng, 1))
        groups = math.ceil(NG[ng] * nprb / 8 - 1e-9)
        for cfi in (1, 2, 3):
            nregs = nprb * (2 + (3 if cfi >= 2 else 0) + (3 if cfi >= 3 else 0))
            want_cce = (nregs - 4 - cfi * groups) // 9                      # one PHICH REG per group in each symbol of the control region
            n, idx, pc = product_map(cellp, ng, cfi, ext=1)
            assert n == want_cce == O.lteo_nof_cce(o.h, cfi), (cellp, ng, cfi, n, want_cce)
            oidx = np.zeros(88 * 36, np.uint16)
            opc = np.zeros(16, np.uint16)
            assert O.lteo_pdcch_re_index(o.h, cfi, ltelib.ptr(oidx), ltelib.ptr(opc)) == n
            assert np.array_equal(idx, oidx[:36 * n]) and np.array_equal(pc, opc), (cellp, ng, cfi)
            assert len(set(idx.tolist())) == len(idx) and not set(idx.tolist()) & set(pc.tolist())
            if cfi == 3 and groups:
                assert not np.array_equal(idx, product_map(cellp, ng, cfi, ext=0)[1])      # a different grid than with the normal duration


def test_transmitter_to_oracle_loop_with_extended_phich(infra):
    from helpers import make_capture
    cell = Cell(50, 2, 21, 2, 0, 2, 1)                                       # Ng = 1, extended duration, CFI 3 (the duration is the lower bound of the CFI)
    sim, iq, tti, truths, payloads = make_capture(cell, 5, seed=40, cfi=3, nof_ues=6, dl_min=3, dl_max=5, tm=2, mcs_min=4, mcs_max=12, snr_db=25.0)
    o = Oracle(cell)
    found = total = 0
    for sf in range(5):
        sym = o.ofdm(iq[sf])
        ce, res = o.chest(int(tti[sf]) % 10, sym)
        cfi, corr = o.pcfich(int(tti[sf]) % 10, sym, ce)
        assert cfi == 3
        llr = o.pdcch_llr(int(tti[sf]) % 10, cfi, sym, ce)
        for i in range(truths[sf].nof_dci):
            d = truths[sf].dci[i]
            total += 1
            r, bits, crc = o.dci_decode(llr[72 * d.ncce:72 * (d.ncce + (1 << d.L))], d.nbits)
            found += r == 0 and crc == d.rnti and np.array_equal(bits, np.frombuffer(bytes(d.bits), np.uint8)[:d.nbits])
    assert total >= 12 and found == total, (total, found)
    L = capi.load_library()
    assert L.ltephy_search_create_cell_ng(50, 2, 21, 2, 2 | (1 << 8), 5) and L.ltephy_search_create_cell_ng(50, 2, 21, 2, 2 | (2 << 8), 5) is None


@pytest.mark.parametrize("ng,ext,cfi", [(2, 0, 2), (3, 0, 3), (1, 1, 3)])
def test_reference_walk_through_the_compat_layer_with_the_mibs_phich_configuration(infra, ng, ext, cfi):
    """tier 2 with a cell as the reference's live mode gets it from the MIB: the reference's own DCISearch.cc runs on srsran_ue_dl_set_cell(cell with
    phich_resources / phich_length) of the compat layer and walks the locations of THAT CCE grid; the product search created with the same configuration must
    accept the same DCIs on the same candidate tables (synthetic eNB with this configuration -> oracle -> tables)"""
    import os as _os
    from test_reference_code import RefWalk, reflib, phase_a_oracle, REF_SO
    from test_host_search import host_geometry
    if not (_os.path.exists(REF_SO) or _os.path.isdir("/root/reference")):
        pytest.skip("reference sources not available")
    cell = Cell(50, 2, 77, 2, 0, ng, ext)
    R = reflib()
    R.refwalk_create_phich.restype = C.c_void_p
    R.refwalk_create_phich.argtypes = [C.c_uint32] * 7
    ref = RefWalk.__new__(RefWalk)
    ref.L = R
    ref.h = R.refwalk_create_phich(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, 5, ng, ext)
    assert ref.h
    srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, phich_resources=ng | (ext << 8))
    s = Sim(cell=cell, seed=50 + ng, cfi=cfi, nof_ues=6, dl_min=3, dl_max=5, ul_min=1, ul_max=2, tm=2, mcs_min=4, mcs_max=12, snr_db=25.0, fixed_L=2)
    o = Oracle(cell)
    geo = host_geometry(cell)
    total = 0
    for tti in range(24):
        info, T, llr, tr = phase_a_oracle(s, o, geo, tti)
        assert info.nof_cce == product_map((50, 2, 77, 2), ng, info.cfi, ext)[0]
        want = ref.subframe(info, T, llr)
        got = srch.subframe(info, T)
        for is_ul in (False, True):
            a = [d for d in got if (d["format"] == 0) == is_ul]
            b = [r for r in want if (r.format == 0) == is_ul]
            assert [(int(d["rnti"]), int(d["format"]), int(d["L"]), int(d["ncce"]), int(d["histogram_value"])) for d in a] == \
                   [(r.rnti, r.format, r.L, r.ncce, r.histval) for r in b], (ng, ext, tti, is_ul)
        total += len(got)
    rs, ps = ref.stats(), srch.stats()
    assert (rs.nof_decoded_locations, rs.nof_cce, rs.nof_locations) == (ps.nof_decoded_locations, ps.nof_cce, ps.nof_locations)
    assert total >= 24
    ref.close()
