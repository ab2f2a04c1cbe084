"""Shared helpers for the parity tests: run the CPU oracle over a capture, convert grants."""
import numpy as np
import ltelib
from ltelib import Sim
from ltesniffer_b200 import capi


def make_capture(cell, n, **simkw):
    s = Sim(cell=cell, **simkw)
    iq = np.zeros((n, cell.nof_rx, s.sf_len), np.complex64)
    truths, payloads = [], []
    tti0 = simkw.get("tti0", 0)
    for i in range(n):
        x, tr, pl = s.subframe(tti0 + i)
        iq[i] = x
        truths.append(tr)
        payloads.append(pl)
    return s, iq, np.arange(tti0, tti0 + n, dtype=np.uint32), truths, payloads


def oracle_frontend(o, iq, tti):
    """-> list of dicts with sym, ce, res, cfi, corr, llr, rb_power, cce_power per subframe"""
    out = []
    for i in range(len(tti)):
        sf_idx = int(tti[i]) % 10
        sym = o.ofdm(iq[i])
        ce, res = o.chest(sf_idx, sym)
        cfi, corr = o.pcfich(sf_idx, sym, ce)
        llr = o.pdcch_llr(sf_idx, cfi, sym, ce)
        ncce = len(llr) // 72
        pw = np.zeros(ncce, np.float32)
        ltelib.oracle().lteo_cce_power(ltelib.ptr(llr), ncce, ltelib.ptr(pw))
        out.append(dict(sym=sym, ce=ce, res=res, cfi=cfi, corr=corr, llr=llr, rb_power=o.rb_power(sym[0]), cce_power=pw))
    return out


def to_phy_grant(sf, rnti, g):
    """ltelib.DlGrant (oracle/sim) -> capi.Grant (product C-ABI)"""
    pg = capi.Grant()
    pg.sf = sf
    pg.rnti = rnti
    pg.tx_scheme = g.tx_scheme
    pg.nof_tb = g.nof_tb
    pg.nof_re = g.nof_re
    pg.pmi = g.pmi
    for s in range(2):
        for prb in range(110):
            if g.prb_mask[s][prb]:
                pg.prb_mask[s][prb >> 5] |= (1 << (prb & 31))
    for t in range(2):
        pg.tb[t].tbs = g.tb[t].tbs
        pg.tb[t].qm = g.tb[t].qm
        pg.tb[t].rv = g.tb[t].rv
        pg.tb[t].enabled = g.tb[t].enabled
    en = [t for t in range(2) if g.tb[t].enabled]
    for k, t in enumerate(en):
        pg.tb[t].cw_idx = (1 - k) if (len(en) == 2 and g.cw_swap) else k
    return pg


def truth_grants(cell, truths, tti, alt=0):
    """ground-truth DL grants of a capture: list of (sf, truth_dci, DlGrant)"""
    out = []
    for sf, tr in enumerate(truths):
        for i in range(tr.nof_dci):
            d = tr.dci[i]
            if d.nof_tb == 0:
                continue
            bits = np.frombuffer(bytes(d.bits), np.uint8)[:d.nbits]
            r, dd, g = ltelib.unpack_and_grant(cell, d.format, d.rnti, bits, int(tti[sf]) % 10, tr.cfi, alt)
            assert r == 0, r
            out.append((sf, d, g))
    return out


def feq(a, b):
    """float arrays equal as values (so -0.0 == +0.0), no NaNs allowed."""
    a = np.asarray(a)
    b = np.asarray(b)
    return a.shape == b.shape and not np.isnan(a).any() and bool(np.all(a == b))


def describe_mismatch(a, b, name):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    bad = np.nonzero(a != b)[0]
    if len(bad) == 0:
        return "%s: identical" % name
    d = np.abs(a.astype(np.complex128) - b.astype(np.complex128))
    return "%s: %d/%d differ, first at %d (%r vs %r), max abs diff %.3g, rel %.3g" % (
        name, len(bad), len(a), bad[0], a[bad[0]], b[bad[0]], d.max(), d.max() / (np.abs(b).max() + 1e-30))
