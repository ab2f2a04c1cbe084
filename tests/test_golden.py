"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from seeded synthetic captures and the
CPU oracle).  CPU: today's oracle must reproduce the committed bytes (the oracle cannot drift silently).  GPU: the CUDA path
must reproduce them through the C-ABI with no oracle in the loop."""
import glob
import os
import numpy as np
import pytest
import ltelib
from ltelib import Cell, Oracle, UlCfg
from helpers import oracle_frontend, feq
from test_host_search import oracle_table, host_geometry, locations
from ltesniffer_b200 import capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DL = sorted(glob.glob(os.path.join(GOLD, "dl_*.npz")))
UL = sorted(glob.glob(os.path.join(GOLD, "ul_*.npz")))


def test_fixtures_present():
    assert len(DL) >= 2 and len(UL) >= 1


def _phy_grant(sf, rnti, prb, num):
    pg = capi.Grant()
    pg.sf, pg.rnti, pg.tx_scheme, pg.nof_tb, pg.nof_re, pg.pmi = int(sf), int(rnti), int(num[0]), int(num[1]), int(num[2]), int(num[3])
    for s in range(2):
        for p in range(110):
            if prb[s, p]:
                pg.prb_mask[s][p >> 5] |= 1 << (p & 31)
    for t in range(2):
        pg.tb[t].enabled, pg.tb[t].qm, pg.tb[t].rv, pg.tb[t].tbs = int(num[4 + 4 * t]), int(num[5 + 4 * t]), int(num[6 + 4 * t]), int(num[7 + 4 * t])
    if pg.tb[0].enabled and pg.tb[1].enabled:      # the fixtures carry no swapped grants: TB 1 -> codeword 0, TB 2 -> codeword 1
        pg.tb[1].cw_idx = 1
    return pg


@pytest.mark.parametrize("path", DL, ids=[os.path.basename(p) for p in DL])
def test_oracle_reproduces_dl_golden(infra, path):
    z = np.load(path)
    cell = Cell(*[int(v) for v in z["cell"]])
    o = Oracle(cell)
    ref = oracle_frontend(o, z["iq"], z["tti"])
    for i in range(len(z["tti"])):
        r = ref[i]
        assert r["cfi"] == int(z["sf%d_cfi" % i])
        assert np.float32(r["res"].snr_db).tobytes() == z["sf%d_snr_db" % i].tobytes()
        assert feq(r["llr"], z["sf%d_llr" % i]) and feq(r["cce_power"], z["sf%d_cce_power" % i])
        nc, Ls = locations(len(r["llr"]) // 72)
        T = oracle_table(o, host_geometry(cell), nc, Ls, r["llr"])[:len(nc)]
        for f in ("valid", "rnti", "bits"):
            assert np.array_equal(T[f], z["sf%d_table" % i][f]), (i, f)
    off = 0
    for k in range(len(z["grant_sf"])):
        sf = int(z["grant_sf"][k])
        g = ltelib.DlGrant.from_buffer_copy(z["grant_raw"][k].tobytes())
        rr, opl, ook = o.pdsch_decode(int(z["tti"][sf]) % 10, ref[sf]["cfi"], int(z["grant_rnti"][k]), g, ref[sf]["sym"], ref[sf]["ce"], 8)
        assert rr == 0
        for t in range(2):
            n = int(z["tb_len"][2 * k + t])
            if n:
                assert ook[t] == int(z["tb_crc"][2 * k + t]) and np.array_equal(opl[t][:n], z["tb_bytes"][off:off + n])
            off += n


@pytest.mark.parametrize("path", UL, ids=[os.path.basename(p) for p in UL])
def test_oracle_reproduces_ul_golden(infra, path):
    z = np.load(path)
    cell = Cell(*[int(v) for v in z["cell"]])
    o = Oracle(cell)
    ucfg = UlCfg(n_dmrs1=int(z["ucfg"][0]), delta_ss=int(z["ucfg"][1]))
    gr = [ltelib.UlGrant(rnti=int(n[0]), qm=int(n[1]), rv=int(n[2]), L_prb=int(n[3]), n_prb=int(n[4]), n_dmrs2=int(n[5]), tbs=int(n[6]),
                         nof_re=144 * int(n[3]), nof_bits=144 * int(n[3]) * int(n[1])) for n in z["grant_num"]]
    sym, ref = ltelib.oracle_ul(o, ucfg, int(z["tti"]), gr, z["iq"])
    assert feq(sym, z["ul_sym"])
    off = 0
    for k, (r, g) in enumerate(zip(ref, gr)):
        n = int(z["tb_len"][k])
        assert r[0] == 0 and r[2] == int(z["tb_crc"][k]) and np.array_equal(r[1][:n], z["tb_bytes"][off:off + n])
        assert np.array([r[3].noise, r[3].rsrp, r[3].snr_db], np.float32).tobytes() == z["chest"][k].tobytes()
        off += n


@pytest.mark.gpu
@pytest.mark.parametrize("path", DL, ids=[os.path.basename(p) for p in DL])
def test_cuda_reproduces_dl_golden(phylib, path):
    z = np.load(path)
    cp = [int(v) for v in z["cell"]]
    n = len(z["tti"])
    phy = capi.LtePhy(cp[0], cp[1], cp[2], cp[3], max_subframes=n)
    phy.submit_iq(z["iq"], z["tti"])
    info, cands = phy.get_phase_a()
    llr = phy.tap(capi.TAP_LLR, (n, capi.LLR_STRIDE), np.float32)
    for i in range(n):
        g = z["sf%d_llr" % i]
        assert info[i].cfi == int(z["sf%d_cfi" % i]) and np.float32(info[i].snr_db).tobytes() == z["sf%d_snr_db" % i].tobytes()
        assert feq(llr[i, :len(g)], g)
        assert feq(np.array(info[i].cce_power[:len(z["sf%d_cce_power" % i])], np.float32), z["sf%d_cce_power" % i])
        T = z["sf%d_table" % i]
        mine = cands[i, :T.shape[0]]
        for f in ("valid", "rnti", "bits"):
            assert np.array_equal(mine[f][:, :T.shape[1]], T[f]), (i, f)
    grants = [_phy_grant(z["grant_sf"][k], z["grant_rnti"][k], z["grant_prb"][k], z["grant_num"][k]) for k in range(len(z["grant_sf"]))]
    phy.submit_grants(grants)
    res, pl = phy.get_phase_b()
    off = 0
    for k in range(2 * len(grants)):
        nby = int(z["tb_len"][k])
        if nby:
            r = res[k]
            assert r.crc == int(z["tb_crc"][k]) and r.payload_len == nby
            assert np.array_equal(pl[r.payload_off:r.payload_off + nby], z["tb_bytes"][off:off + nby]), k
        off += nby
    phy.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", UL, ids=[os.path.basename(p) for p in UL])
def test_cuda_reproduces_ul_golden(phylib, path):
    z = np.load(path)
    cp = [int(v) for v in z["cell"]]
    phy = capi.LtePhy(cp[0], cp[1], cp[2], cp[3], max_subframes=1)
    phy.set_ul_cfg(int(z["ucfg"][0]), int(z["ucfg"][1]))
    grants = [capi.UlGrant(sf=0, rnti=int(n[0]), qm=int(n[1]), rv=int(n[2]), L_prb=int(n[3]), n_prb=int(n[4]), n_dmrs2=int(n[5]), tbs=int(n[6])) for n in z["grant_num"]]
    res, ch, pl = phy.decode_ul(z["iq"][None, :], np.array([int(z["tti"])], np.uint32), grants)
    sym = phy.tap(capi.TAP_UL_SYM, (1, 14 * phy.nsc), np.complex64)
    assert feq(sym[0], z["ul_sym"])
    off = 0
    for k in range(len(grants)):
        nby = int(z["tb_len"][k])
        assert res[k].crc == int(z["tb_crc"][k]) and np.array_equal(pl[res[k].payload_off:res[k].payload_off + nby], z["tb_bytes"][off:off + nby])
        assert np.array([ch[k].noise, ch[k].rsrp, ch[k].snr_db], np.float32).tobytes() == z["chest"][k].tobytes()
        off += nby
    phy.close()
