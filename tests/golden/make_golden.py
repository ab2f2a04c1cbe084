"""Generates the golden fixtures in this directory: small seeded captures from the synthetic eNB / UEs (sim/) together with
the CPU oracle's outputs for them.  The reference ships no vectors of its own (parity unpinned), so these
pin the ORACLE: tests/test_golden.py checks that today's oracle still reproduces the committed bytes (CPU) and that the
CUDA path reproduces them without the oracle in the loop (GPU).  Run from the repo root: python tests/golden/make_golden.py"""
import os
import sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ltelib  # noqa: E402
from ltelib import Cell, Oracle, UlCfg, Sim  # noqa: E402
from helpers import make_capture, oracle_frontend, truth_grants  # noqa: E402
from test_host_search import oracle_table, host_geometry, locations  # noqa: E402

DL_CASES = {
    "dl_5MHz_1p1a_tm1": dict(cell=(25, 1, 150, 1), n=2, kw=dict(seed=41, cfi=2, nof_ues=3, dl_min=1, dl_max=2, tm=1, mcs_min=3, mcs_max=12, snr_db=26.0, tti0=3)),
    "dl_5MHz_2p2a_tm3": dict(cell=(25, 2, 21, 2), n=2, kw=dict(seed=42, cfi=3, nof_ues=4, dl_min=2, dl_max=3, tm=3, mcs_min=8, mcs_max=20, snr_db=27.0, chan_delay=3)),
}
UL_CASES = {"ul_5MHz": dict(cell=(25, 1, 5, 1), seed=43, tti=6, ngr=2, ucfg=(3, 2))}


def grant_array(g):
    """flatten an oracle DlGrant into integers (what the C-ABI grant needs)"""
    prb = np.zeros((2, 110), np.uint8)
    for s in range(2):
        for p in range(110):
            prb[s, p] = g.prb_mask[s][p]
    return prb, np.array([g.tx_scheme, g.nof_tb, g.nof_re, g.pmi, g.tb[0].enabled, g.tb[0].qm, g.tb[0].rv, g.tb[0].tbs,
                          g.tb[1].enabled, g.tb[1].qm, g.tb[1].rv, g.tb[1].tbs], np.int64)


def make_dl(name, c):
    cell = Cell(*c["cell"])
    sim, iq, tti, truths, payloads = make_capture(cell, c["n"], **c["kw"])
    o = Oracle(cell)
    ref = oracle_frontend(o, iq, tti)
    geo = host_geometry(cell)
    out = dict(cell=np.array(c["cell"], np.uint32), iq=iq, tti=tti)
    for i in range(c["n"]):
        r = ref[i]
        nc, Ls = locations(len(r["llr"]) // 72)
        T = oracle_table(o, geo, nc, Ls, r["llr"])
        out["sf%d_cfi" % i] = np.uint32(r["cfi"])
        out["sf%d_snr_db" % i] = np.float32(r["res"].snr_db)
        out["sf%d_llr" % i] = r["llr"]
        out["sf%d_cce_power" % i] = r["cce_power"]
        out["sf%d_table" % i] = T[:len(nc)]
    tg = truth_grants(cell, truths, tti)
    gsf, grnti, gprb, gnum, tbcrc, tbbytes, graw = [], [], [], [], [], [], []
    for sf, d, g in tg:
        rr, opl, ook = o.pdsch_decode(int(tti[sf]) % 10, ref[sf]["cfi"], d.rnti, g, ref[sf]["sym"], ref[sf]["ce"], 8)
        assert rr == 0
        prb, num = grant_array(g)
        gsf.append(sf), grnti.append(d.rnti), gprb.append(prb), gnum.append(num), graw.append(np.frombuffer(bytes(g), np.uint8))
        for t in range(2):
            n = g.tb[t].tbs // 8 if g.tb[t].enabled else 0
            tbcrc.append(ook[t] if n else 0)
            tbbytes.append(np.asarray(opl[t][:n], np.uint8))
    out.update(grant_sf=np.array(gsf, np.uint32), grant_rnti=np.array(grnti, np.uint32), grant_prb=np.array(gprb), grant_num=np.array(gnum), grant_raw=np.array(graw),
               tb_crc=np.array(tbcrc, np.uint8), tb_len=np.array([len(b) for b in tbbytes], np.uint32), tb_bytes=np.concatenate(tbbytes))
    assert out["tb_crc"].sum() > 0
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "grants", len(gsf), "TBs ok", int(out["tb_crc"].sum()), "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def make_ul(name, c):
    cell = Cell(*c["cell"])
    s = Sim(cell=cell, seed=c["seed"], snr_db=24.0, nof_ues=1, chan_delay=2)
    o = Oracle(cell)
    ucfg = UlCfg(n_dmrs1=c["ucfg"][0], delta_ss=c["ucfg"][1])
    rng = np.random.default_rng(c["seed"])
    gr = ltelib.make_ul_grants(cell, rng, c["ngr"], table=1)
    x, pl, off = ltelib.sim_ul_subframe(s, c["tti"], ucfg, gr)
    sym, ref = ltelib.oracle_ul(o, ucfg, c["tti"], gr, x)
    num = np.array([[g.rnti, g.qm, g.rv, g.L_prb, g.n_prb, g.n_dmrs2, g.tbs] for g in gr], np.int64)
    crc = np.array([r[2] for r in ref], np.uint8)
    chest = np.array([[r[3].noise, r[3].rsrp, r[3].snr_db] for r in ref], np.float32)
    tb = [np.asarray(r[1][:g.tbs // 8], np.uint8) for r, g in zip(ref, gr)]
    assert crc.sum() > 0
    np.savez_compressed(os.path.join(HERE, name + ".npz"), cell=np.array(c["cell"], np.uint32), iq=x, tti=np.uint32(c["tti"]), ucfg=np.array(c["ucfg"], np.uint32),
                        ul_sym=sym, grant_num=num, tb_crc=crc, chest=chest, tb_len=np.array([len(b) for b in tb], np.uint32), tb_bytes=np.concatenate(tb))
    print(name, "grants", len(gr), "TBs ok", int(crc.sum()), "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    for k, v in DL_CASES.items():
        make_dl(k, v)
    for k, v in UL_CASES.items():
        make_ul(k, v)
