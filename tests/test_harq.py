"""HARQ soft-combining mode (reference src/src/HARQ.cc:60-188, src/src/DL_Sniffer_PDSCH.cc:942-1018): the host bookkeeping and, on the CPU,
the combining gain of the oracle's soft buffers on retransmitted transport blocks (rv 0 then rv 2) at an SNR where one transmission is not enough."""
import numpy as np
import pytest
import ltelib
from ltelib import Cell
from helpers import make_capture, truth_grants, oracle_frontend
from ltesniffer_b200 import capi

HARQ_CELL = (25, 1, 77, 1)
HARQ_KW = dict(seed=11, cfi=2, nof_ues=3, dl_min=2, dl_max=3, tm=1, mcs_min=16, mcs_max=16, snr_db=8.0, full_band=1, harq_retx=1)


def test_harq_bookkeeping_follows_the_reference(phylib):
    q = capi.Harq(max_rnti=2)
    # unknown RNTI, free entity: new transmission (HARQ.cc:91-95); slot = (entity * 8 + pid) * 2 + tb
    r, s0 = q.classify(0x1234, 3, 0, 1, 5000, 100)
    assert r == capi.HARQ_NEW_TX
    q.update(0x1234, 3, 0, 1, 0, 5000, 100, False)
    # 8 ms later, same NDI and TBS, not decoded: retransmission into the same slot
    r, s1 = q.classify(0x1234, 3, 0, 1, 5000, 108)
    assert (r, s1) == (capi.HARQ_RE_TX, s0)
    q.update(0x1234, 3, 0, 1, 2, 5000, 108, True)
    # decoded meanwhile: the next repetition is skipped
    assert q.classify(0x1234, 3, 0, 1, 5000, 116)[0] == capi.HARQ_DECODED
    # any other spacing than 8 ms, a toggled NDI or another TBS: new transmission
    assert q.classify(0x1234, 3, 0, 1, 5000, 117)[0] == capi.HARQ_NEW_TX
    assert q.classify(0x1234, 3, 0, 0, 5000, 116)[0] == capi.HARQ_NEW_TX
    assert q.classify(0x1234, 3, 0, 1, 4000, 116)[0] == capi.HARQ_NEW_TX
    # the other TB and other processes have their own slots; tti wraps at 10240
    assert q.classify(0x1234, 3, 1, 1, 5000, 116)[1] == s0 + 1
    assert q.classify(0x1234, 4, 0, 1, 5000, 116)[1] == s0 + 2
    q.update(0x1234, 5, 0, 0, 0, 800, 10236, False)
    assert q.classify(0x1234, 5, 0, 0, 800, 4)[0] == capi.HARQ_RE_TX
    # two entities only: the third RNTI decodes without a store
    assert q.classify(0x2222, 0, 0, 0, 100, 7)[0] == capi.HARQ_NEW_TX
    assert q.classify(0x3333, 0, 0, 0, 100, 7)[0] == capi.HARQ_FULL_BUFFER
    # a process that was never updated counts as "first transmission"
    assert q.classify(0x2222, 1, 0, 0, 100, 15)[0] == capi.HARQ_NEW_TX
    q.close()


def harq_walk(cell, iq, tti, truths, decode):
    """drives the HARQ bookkeeping over the ground-truth C-RNTI grants of a capture in subframe order; decode(sf, d, dd, g, ops, slots) -> [crc per TB]"""
    q = capi.Harq()
    out = []
    for sf, tr in enumerate(truths):
        for i in range(tr.nof_dci):
            d = tr.dci[i]
            if d.nof_tb == 0:
                continue
            bits = np.frombuffer(bytes(d.bits), np.uint8)[:d.nbits]
            r, dd, g = ltelib.unpack_and_grant(cell, d.format, d.rnti, bits, int(tti[sf]) % 10, tr.cfi, 0)
            assert r == 0
            ops, slots = [capi.HARQ_NONE] * 2, [0] * 2
            for t in range(2):
                if not g.tb[t].enabled:
                    continue
                st, slot = q.classify(d.rnti, dd.pid, t, dd.ndi[t], g.tb[t].tbs, int(tti[sf]))
                assert st in (capi.HARQ_NEW_TX, capi.HARQ_RE_TX, capi.HARQ_DECODED)
                ops[t] = {capi.HARQ_NEW_TX: capi.HARQ_NEW, capi.HARQ_RE_TX: capi.HARQ_RETX, capi.HARQ_DECODED: -1}[st]
                slots[t] = slot
            crc = decode(sf, d, dd, g, ops, slots)
            for t in range(2):
                if g.tb[t].enabled and ops[t] in (capi.HARQ_NEW, capi.HARQ_RETX):
                    q.update(d.rnti, dd.pid, t, dd.ndi[t], dd.rv[t], g.tb[t].tbs, int(tti[sf]), crc[t])
            out.append((sf, d.rnti, ops, crc))
    q.close()
    return out


def oracle_harq_decoder(o, fe, tti, store):
    def decode(sf, d, dd, g, ops, slots):
        soft, comb = [None, None], [0, 0]
        for t in range(2):
            if ops[t] in (capi.HARQ_NEW, capi.HARQ_RETX):
                soft[t] = store.setdefault(slots[t], np.zeros(16 * ltelib.HARQ_CB_STRIDE, np.int16))
                comb[t] = 1 if ops[t] == capi.HARQ_RETX else 0
        if all(op == -1 or not g.tb[t].enabled for t, op in enumerate(ops)):
            return [1, 1]      # already decoded: skipped
        r, pl, ok = o.pdsch_decode_harq(int(tti[sf]) % 10, fe[sf]["cfi"], d.rnti, g, fe[sf]["sym"], fe[sf]["ce"], soft, comb)
        assert r == 0
        decode.payloads[(sf, d.rnti)] = pl
        return ok
    decode.payloads = {}
    return decode


def test_oracle_soft_combining_gain(infra):
    cell = Cell(*HARQ_CELL)
    sim, iq, tti, truths, payloads = make_capture(cell, 16, **HARQ_KW)
    # the second half repeats the first half's DCIs with rv 2 and the same payload
    for sf in range(8):
        a, b = truths[sf], truths[sf + 8]
        assert a.nof_dci == b.nof_dci and a.nof_dci >= 2
        for i in range(a.nof_dci):
            assert (a.dci[i].rnti, a.dci[i].tbs[0], a.dci[i].rv[0], b.dci[i].rv[0]) == (b.dci[i].rnti, b.dci[i].tbs[0], 0, 2)
        assert np.array_equal(payloads[sf], payloads[sf + 8])
    o = ltelib.Oracle(cell)
    fe = oracle_frontend(o, iq, tti)
    res = harq_walk(cell, iq, tti, truths, oracle_harq_decoder(o, fe, tti, {}))
    first = [crc[0] for sf, _, ops, crc in res if sf < 8]
    second = [(ops[0], crc[0]) for sf, _, ops, crc in res if sf >= 8]
    assert sum(first) <= len(first) // 3                       # one transmission at this SNR mostly fails ...
    comb = [c for op, c in second if op == capi.HARQ_RETX]
    assert len(comb) >= len(first) - sum(first) and sum(comb) >= len(comb) - 1      # ... two combined decode
    assert all(op == -1 for (op, c), f in zip(second, first) if f)                  # decoded first time: the repetition is skipped
    # the same retransmissions decoded on their own (no store) fail like the first ones
    alone = 0
    for sf in range(8, 16):
        for s2, d, g in [x for x in truth_grants(cell, truths, tti) if x[0] == sf]:
            r, pl, ok = o.pdsch_decode(int(tti[sf]) % 10, fe[sf]["cfi"], d.rnti, g, fe[sf]["sym"], fe[sf]["ce"])
            alone += ok[0]
    assert alone <= sum(comb) // 2


@pytest.mark.gpu
@pytest.mark.parametrize("split", [8, 16])
def test_gpu_harq_store_matches_oracle(infra, phylib, split):
    """product (rate-dematch accumulators kept in the HBM store across transmissions) == oracle with per-slot soft buffers: CRC flags and payload
    bytes of every transport block.  split = 8: first transmissions and retransmissions in different batches; split = 16: one batch holds both,
    so every slot is used twice inside it (second rate-dematch launch)."""
    from helpers import to_phy_grant
    cell = Cell(*HARQ_CELL)
    sim, iq, tti, truths, payloads = make_capture(cell, 16, **HARQ_KW)
    o = ltelib.Oracle(cell)
    fe = oracle_frontend(o, iq, tti)
    odec = oracle_harq_decoder(o, fe, tti, {})
    ref = harq_walk(cell, iq, tti, truths, odec)
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=16, turbo_max_iter=8)
    phy.harq_reserve(150 * 16)
    got = []
    q = capi.Harq()
    for b0 in range(0, 16, split):
        sel = slice(b0, b0 + split)
        phy.submit_iq(iq[sel], tti[sel])
        phy.get_phase_a()
        grants, meta = [], []
        for sf in range(b0, b0 + split):
            tr = truths[sf]
            for i in range(tr.nof_dci):
                d = tr.dci[i]
                if d.nof_tb == 0:
                    continue
                bits = np.frombuffer(bytes(d.bits), np.uint8)[:d.nbits]
                r, dd, g = ltelib.unpack_and_grant(cell, d.format, d.rnti, bits, int(tti[sf]) % 10, tr.cfi, 0)
                pg = to_phy_grant(sf - b0, d.rnti, g)
                st, slot = q.classify(d.rnti, dd.pid, 0, dd.ndi[0], g.tb[0].tbs, int(tti[sf]))
                if split == 16 and st == capi.HARQ_NEW_TX and sf >= 8:
                    st = capi.HARQ_RE_TX          # inside one batch the first decode is not known yet: combine (never skip), see DESIGN.md
                op = {capi.HARQ_NEW_TX: capi.HARQ_NEW, capi.HARQ_RE_TX: capi.HARQ_RETX}.get(st, -1)
                if op == -1:
                    meta.append((sf, d.rnti, dd, g, None))
                    continue
                pg.tb[0].harq_op, pg.tb[0].harq_slot = op, slot
                if split == 16:
                    q.update(d.rnti, dd.pid, 0, dd.ndi[0], dd.rv[0], g.tb[0].tbs, int(tti[sf]), False)
                grants.append(pg)
                meta.append((sf, d.rnti, dd, g, len(grants) - 1))
        phy.submit_grants(grants)
        res, pl = phy.get_phase_b()
        for sf, rnti, dd, g, gi in meta:
            if gi is None:
                got.append((sf, rnti, -1, 1, None))
                continue
            r = res[2 * gi]
            got.append((sf, rnti, grants[gi].tb[0].harq_op, r.crc, bytes(pl[r.payload_off:r.payload_off + r.payload_len])))
            if split == 8:
                q.update(rnti, dd.pid, 0, dd.ndi[0], dd.rv[0], g.tb[0].tbs, int(tti[sf]), r.crc)
    q.close()
    phy.close()
    assert len(got) == len(ref)
    ncomb = 0
    for (sf, rnti, op, crc, pl), (rsf, rrnti, rops, rcrc) in zip(got, ref):
        assert (sf, rnti) == (rsf, rrnti)
        if split == 8:
            assert op == rops[0] and crc == rcrc[0], (sf, hex(rnti), op, rops, crc, rcrc)
        elif rops[0] != -1:
            assert crc == rcrc[0], (sf, hex(rnti), crc, rcrc)       # blocks the oracle skipped as decoded are decoded again here (and pass)
        else:
            assert crc == 1
        if pl is not None and crc:
            opl = odec.payloads.get((sf, rnti))
            if opl is not None and rops[0] != -1:
                assert pl == bytes(opl[0][:len(pl)])
        ncomb += op == capi.HARQ_RETX and crc
    assert ncomb >= 15


def test_harq_prepare_grant_fills_the_grant(phylib):
    """ltephy_harq_prepare_grant: classification + the grant fields the rate-dematcher reads; a decoded single-TB grant is disabled as the reference does
    (pdsch_cfg->grant.tb[i].enabled = false, DL_Sniffer_PDSCH.cc:970-972)"""
    import ctypes as C
    L = capi.load_library()

    class Fields(C.Structure):
        _fields_ = [("rnti", C.c_uint16), ("format", C.c_uint8), ("alloc_type", C.c_uint8), ("mcs", C.c_uint8 * 2), ("rv", C.c_uint8 * 2), ("ndi", C.c_uint8 * 2),
                    ("harq_pid", C.c_uint8), ("tpc", C.c_uint8), ("tb_cw_swap", C.c_uint8), ("pinfo", C.c_uint8), ("nof_prb", C.c_uint32)]
    q = capi.Harq(max_rnti=4)
    L.ltephy_harq_prepare_grant.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    f = Fields(rnti=0x4321, harq_pid=5)
    f.ndi[0] = 1
    g = capi.Grant(rnti=0x4321)
    g.tb[0].enabled, g.tb[0].tbs = 1, 2216
    st = (C.c_int * 2)()
    assert L.ltephy_harq_prepare_grant(q.h, C.byref(f), 200, C.byref(g), st) == 0
    assert (st[0], st[1], g.tb[0].harq_op) == (capi.HARQ_NEW_TX, -1, capi.HARQ_NEW)
    slot = g.tb[0].harq_slot
    q.update(0x4321, 5, 0, 1, 0, 2216, 200, False)
    g.tb[0].harq_op = 0
    assert L.ltephy_harq_prepare_grant(q.h, C.byref(f), 208, C.byref(g), st) == 0
    assert (st[0], g.tb[0].harq_op, g.tb[0].harq_slot, g.tb[0].enabled) == (capi.HARQ_RE_TX, capi.HARQ_RETX, slot, 1)
    q.update(0x4321, 5, 0, 1, 2, 2216, 208, True)
    assert L.ltephy_harq_prepare_grant(q.h, C.byref(f), 216, C.byref(g), st) == 0
    assert (st[0], g.tb[0].enabled, g.tb[0].harq_op) == (capi.HARQ_DECODED, 0, capi.HARQ_NONE)
    q.close()


def test_reserved_mcs_block_is_sized_from_the_harq_process(phylib, infra):
    """An adaptive retransmission with a reserved MCS (29-31) carries no size: ltephy_dci_to_grant leaves tbs = 0, and in HARQ mode the reference gives
    the block the size of its process' last transmission before the "tbs > 0" skip rule (DCICollection::addCandidate, src/src/DCICollection.cc:236-252;
    HARQ::getlastTbs, src/src/HARQ.cc:262-274).  Here: ltephy_search_keep_reserved_mcs lets the grant through ltephy_grants_from_dcis,
    ltephy_harq_prepare_grant sizes and classifies it; without a recorded transmission it stays at 0 and is not decoded."""
    import ctypes as C
    S = infra.sim()
    L = capi.load_library()
    capi._bind_search(L)
    cell = Cell(25, 1, 77, 1)
    srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
    rng = np.random.default_rng(4)
    nb = S.lte_dci_sizeof(C.byref(cell), 1)
    rnti, tti, cfi = 0x2345, 508, 2
    row = np.zeros(1, capi.DCI_DTYPE)
    for _ in range(20000):                                    # a format-1 payload whose MCS field reads 29 and which is a valid grant otherwise
        bits = rng.integers(0, 2, nb).astype(np.uint8)
        v = 0
        for k, b in enumerate(bits):
            v |= int(b) << (63 - k)
        row["sf"], row["rnti"], row["format"], row["nof_bits"], row["bits"] = 0, rnti, 1, nb, v
        rc, g, f = srch.dci_to_grant(row[0], tti % 10, cfi, 0)
        if rc == 0 and f.mcs[0] == 29:
            break
    assert rc == 0 and f.mcs[0] == 29 and g.tb[0].enabled and g.tb[0].tbs == 0 and g.tb[0].qm == 2
    info = (capi.SfInfo * 1)()
    info[0].tti, info[0].cfi = tti, cfi
    grants = (capi.Grant * 4)()
    gidx = np.zeros(4, np.uint32)
    ng = C.c_uint32(0)

    def build():
        assert L.ltephy_grants_from_dcis(srch.h, info, row.ctypes.data_as(C.c_void_p), 1, 1, 0, grants, gidx.ctypes.data_as(C.c_void_p), 4, C.byref(ng)) == 0
        return ng.value
    assert build() == 0                                       # decode_dl_mode's skip rule (DL_Sniffer_PDSCH.cc:887)
    L.ltephy_search_keep_reserved_mcs(srch.h, 1)
    assert build() == 1 and grants[0].tb[0].tbs == 0 and grants[0].rnti == rnti
    L.ltephy_harq_prepare_grant.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    q = capi.Harq(max_rnti=4)
    st = (C.c_int * 2)()
    gr = capi.Grant.from_buffer_copy(grants[0])
    assert L.ltephy_harq_prepare_grant(q.h, C.byref(f), tti, C.byref(gr), st) == 0
    assert (st[0], gr.tb[0].tbs, gr.tb[0].harq_op, gr.tb[0].enabled) == (-1, 0, capi.HARQ_NONE, 1)      # nothing known about this process
    assert q.classify(rnti, f.harq_pid, 0, f.ndi[0], 2216, tti - 8)[0] == capi.HARQ_NEW_TX               # the first transmission, 8 ms earlier, failed
    q.update(rnti, f.harq_pid, 0, f.ndi[0], 0, 2216, tti - 8, False)
    assert q.last_tbs(rnti, f.harq_pid, 0) == 2216 and q.last_tbs(rnti, (f.harq_pid + 1) % 8, 0) == 0 and q.last_tbs(0x999, 0, 0) == 0
    gr = capi.Grant.from_buffer_copy(grants[0])
    assert L.ltephy_harq_prepare_grant(q.h, C.byref(f), tti, C.byref(gr), st) == 0
    assert (st[0], gr.tb[0].tbs, gr.tb[0].harq_op, gr.tb[0].qm) == (capi.HARQ_RE_TX, 2216, capi.HARQ_RETX, 2)
    q.close()
