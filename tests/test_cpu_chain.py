"""CPU tests of the test infrastructure itself: synthetic eNB -> oracle receiver loop (ground truth is
CRC-self-validating), DCI size known answers, segmentation/rate-matching properties."""
import ctypes as C
import numpy as np
import pytest
import ltelib
from ltelib import Cell, Sim, Oracle, FORMATS


def test_dci_sizes_known_answers(infra):
    S = infra.sim()
    c = Cell(100, 2, 0, 1)   # 0/1A 28, 1 39, 1B 30, 1C 15, 1D 30, 2 51, 2A 48
    assert [S.lte_dci_sizeof(C.byref(c), f) for f in range(9)] == [28, 39, 28, 30, 15, 30, 51, 48, 48]
    c = Cell(50, 2, 0, 1)
    assert [S.lte_dci_sizeof(C.byref(c), f) for f in (0, 1, 2, 4, 6, 7)] == [27, 31, 27, 13, 43, 41]
    c = Cell(25, 1, 0, 1)
    assert [S.lte_dci_sizeof(C.byref(c), f) for f in (0, 1, 2, 4)] == [25, 27, 25, 12]


def test_control_region_sizes(infra):
    o = Oracle(Cell(100, 2, 1, 1))
    assert [infra.oracle().lteo_nof_cce(o.h, cfi) for cfi in (1, 2, 3)] == [20, 54, 87]


@pytest.mark.parametrize("tbs,rv,qm", [(40, 0, 2), (2216, 0, 4), (6200, 2, 6), (36696, 0, 6), (75376, 0, 6), (75376, 1, 6), (97896, 0, 8)])
def test_dlsch_encode_decode_roundtrip(infra, tbs, rv, qm):
    """encode -> (noise free) soft bits -> rate-dematch + turbo -> CRC ok and identical bytes"""
    S, O = infra.sim(), infra.oracle()
    rng = np.random.default_rng(tbs + rv)
    pl = rng.integers(0, 256, tbs // 8).astype(np.uint8)
    G = int((tbs + 24) * (1.6 if rv == 0 else 2.4)) // (2 * qm) * 2 * qm
    e = np.zeros(G, np.uint8)
    assert S.lte_sim_dlsch_encode(ltelib.ptr(pl), tbs, rv, G, qm, 1, ltelib.ptr(e)) == 0
    llr = ((2 * e.astype(np.int16) - 1) * 120).astype(np.int16)
    out = np.zeros(tbs // 8 + 8, np.uint8)
    it = np.zeros(32, np.uint32)
    assert O.lteo_dlsch_decode(ltelib.ptr(llr), G, tbs, rv, qm, 1, 4, 1, ltelib.ptr(out), ltelib.ptr(it)) == 1
    assert np.array_equal(out[:tbs // 8], pl)
    # erase 15 % of the soft bits: still decodable (encode -> erase -> decode property)
    llr2 = llr.copy()
    llr2[rng.random(G) < 0.15] = 0
    assert O.lteo_dlsch_decode(ltelib.ptr(llr2), G, tbs, rv, qm, 1, 8, 1, ltelib.ptr(out), ltelib.ptr(it)) == 1
    assert np.array_equal(out[:tbs // 8], pl)


def test_pdcch_encode_decode_all_formats(infra):
    S = infra.sim()
    cell = Cell(100, 2, 5, 1)
    o = Oracle(cell)
    rng = np.random.default_rng(1)
    for f in range(9):
        nb = S.lte_dci_sizeof(C.byref(cell), f)
        for L in range(4):
            if 3 * (nb + 16) > (72 << L) * 3:
                continue
            b = rng.integers(0, 2, nb).astype(np.uint8)
            rnti = int(rng.integers(1, 65535))
            e = np.zeros(72 << L, np.uint8)
            S.lte_sim_pdcch_encode(ltelib.ptr(b), nb, rnti, L, ltelib.ptr(e))
            llr = (2.0 * e - 1.0 + 0.25 * rng.standard_normal(72 << L)).astype(np.float32)
            if nb + 16 > 0.9 * (72 << L):
                continue
            r, bits, crc = o.dci_decode(llr, nb)
            assert r == 0 and crc == rnti and np.array_equal(bits, b), (FORMATS[f], L)
    # all-zero input is not decoded (falcon_pdcch.c:141: mean > 0)
    assert o.dci_decode(np.zeros(72, np.float32), 28)[0] == -1


CAPS = {
    "cfg1_10sf_1rnti_tm1_qpsk": (Cell(100, 1, 1, 1), 10, dict(seed=1, cfi=2, nof_ues=1, tm=1, mcs_min=5, mcs_max=5, snr_db=30.0, fixed_L=2, si_period=5)),
    "tm3_2x2_64qam": (Cell(100, 2, 7, 2), 2, dict(seed=2, cfi=3, nof_ues=150, dl_min=8, dl_max=12, tm=3, mcs_min=17, mcs_max=24, snr_db=28.0, full_band=1)),
    "mix_50prb_delay": (Cell(50, 2, 301, 2), 3, dict(seed=3, cfi=3, nof_ues=20, dl_min=3, dl_max=5, ul_min=1, ul_max=2, tm=13, mcs_min=0, mcs_max=20, snr_db=22.0, chan_delay=6)),
    "tm4_2x2_256qam_alt_table": (Cell(50, 2, 11, 2), 3, dict(seed=6, cfi=2, nof_ues=8, dl_min=2, dl_max=4, tm=4, mcs_min=4, mcs_max=22, snr_db=33.0, alt_table=1)),
    "tm3_cw_swap": (Cell(50, 2, 21, 2), 4, dict(seed=8, cfi=2, nof_ues=8, dl_min=2, dl_max=4, tm=3, mcs_min=6, mcs_max=24, snr_db=29.0, tb_swap=1)),
    "tm4_cw_swap": (Cell(50, 2, 11, 2), 4, dict(seed=9, cfi=2, nof_ues=8, dl_min=2, dl_max=4, tm=4, mcs_min=4, mcs_max=20, snr_db=33.0, tb_swap=1)),
    "sf0_sf5_sync_re_exclusion_25prb": (Cell(25, 2, 77, 1), 6, dict(seed=4, cfi=2, nof_ues=3, dl_min=1, dl_max=2, tm=1, mcs_min=4, mcs_max=10, snr_db=26.0, full_band=1)),
}


@pytest.mark.parametrize("name", list(CAPS))
def test_sim_to_oracle_ground_truth(infra, name):
    """every transmitted DCI and transport block is recovered by the oracle receiver"""
    cell, n, kw = CAPS[name]
    s, o = Sim(cell=cell, **kw), Oracle(cell)
    ndci = ntb = nswap = 0
    for tti in range(n):
        iq, tr, pl = s.subframe(tti)
        sym = o.ofdm(iq)
        ce, res = o.chest(tti % 10, sym)
        cfi, _ = o.pcfich(tti % 10, sym, ce)
        assert cfi == tr.cfi
        assert abs(res.snr_db - kw["snr_db"]) < 6.0
        llr = o.pdcch_llr(tti % 10, cfi, sym, ce)
        for i in range(tr.nof_dci):
            d = tr.dci[i]
            r, bits, crc = o.dci_decode(llr[72 * d.ncce:72 * d.ncce + (72 << d.L)], d.nbits)
            assert r == 0 and crc == d.rnti and np.array_equal(bits, np.frombuffer(bytes(d.bits), np.uint8)[:d.nbits])
            ndci += 1
            if d.nof_tb == 0:
                continue
            r, dd, g = ltelib.unpack_and_grant(cell, d.format, crc, bits, tti % 10, cfi, kw.get("alt_table", 0))
            assert r == 0 and g.nof_re == d.nof_re
            nswap += g.cw_swap
            r, pay, ok = o.pdsch_decode(tti % 10, cfi, crc, g, sym, ce, 8)
            assert r == 0
            for t in range(2):
                if g.tb[t].enabled:
                    nby = g.tb[t].tbs // 8
                    assert ok[t] and np.array_equal(pay[t][:nby], pl[d.payload_off[t]:d.payload_off[t] + nby]), (name, tti, hex(d.rnti), t)
                    ntb += 1
    assert ndci >= n and ntb >= 1
    if kw.get("tb_swap"):
        assert nswap >= 2          # the swap flag was exercised (TB 1 on codeword 1, TB 2 on codeword 0) and both TBs still decode


def test_pusch_all_mcs_tables_roundtrip(infra):
    """synthetic UEs -> oracle PUSCH receiver for the three MCS interpretations PUSCH_Decoder::decode tries
    (src/src/UL_Sniffer_PUSCH.cc:498-521): 16QAM cap, 64QAM table, 256QAM table (36.213 Table 8.6.1-3, Qm up to 8)"""
    from ltelib import UlCfg
    cell = Cell(50, 1, 17, 1)
    o = ltelib.Oracle(cell)
    ucfg = UlCfg(n_dmrs1=3, delta_ss=2)
    for table in (0, 1, 2):
        s = ltelib.Sim(cell=cell, seed=5 + table, snr_db=36.0, nof_ues=1, chan_delay=2)
        rng = np.random.default_rng(3 + table)
        tot = ok = 0
        qms = set()
        for tti in range(4, 8):
            gr = ltelib.make_ul_grants(cell, rng, 4, table=table)
            x, pl, off = ltelib.sim_ul_subframe(s, tti, ucfg, gr)
            sym, ref = ltelib.oracle_ul(o, ucfg, tti, gr, x)
            for g, (r, opl, crc, ch, _), of in zip(gr, ref, off):
                assert r == 0
                tot += 1
                ok += crc
                qms.add(g.qm)
                if crc:
                    assert np.array_equal(opl[:g.tbs // 8], pl[of:of + g.tbs // 8])
        assert ok == tot, (table, ok, tot)
        assert max(qms) == (4, 6, 8)[table]


def _ul_roundtrip(cell, ucfg, seed, mutate, n_tti=3, n_gr=3, snr=34.0, min_prb=3, table=1):
    """sim -> oracle PUSCH for grants altered by mutate(g, rng); returns [(grant, crc, chest)]"""
    o = ltelib.Oracle(cell)
    s = ltelib.Sim(cell=cell, seed=seed, snr_db=snr, nof_ues=1)
    rng = np.random.default_rng(seed)
    out = []
    for tti in range(2, 2 + n_tti):
        gr = ltelib.make_ul_grants(cell, rng, n_gr, table=table, min_prb=min_prb)
        for g in gr:
            mutate(g, rng)
        x, pl, off = ltelib.sim_ul_subframe(s, tti, ucfg, gr)
        sym, ref = ltelib.oracle_ul(o, ucfg, tti, gr, x)
        for g, (r, opl, crc, ch, _), of in zip(gr, ref, off):
            assert r == 0
            if crc:
                assert np.array_equal(opl[:g.tbs // 8], pl[of:of + g.tbs // 8])
            out.append((g, crc, ch))
    return out


def test_pusch_uci_multiplexing_roundtrip(infra):
    """HARQ-ACK, RI and CQI multiplexed with the data (36.212 5.2.2.6-8, what PUSCH_Decoder::decode configures at src/src/UL_Sniffer_PUSCH.cc:429-450):
    the UL-SCH still decodes when the receiver takes the CQI / RI symbols out and erases the ACK ones -- and does NOT when it is told nothing"""
    from ltelib import UlCfg
    cell = Cell(50, 1, 23, 1)
    ucfg = UlCfg(n_dmrs1=4, delta_ss=7)
    combos = [dict(nof_ack=1, I_offset_ack=9), dict(nof_ack=2, I_offset_ack=10, ri_len=1, I_offset_ri=8), dict(cqi_len=4, I_offset_cqi=8),
              dict(cqi_len=30, I_offset_cqi=6, ri_len=1, I_offset_ri=5, nof_ack=2, I_offset_ack=5)]
    for ci, cmb in enumerate(combos):
        def mut(g, rng):
            for k, v in cmb.items():
                setattr(g, k, v)
        res = _ul_roundtrip(cell, ucfg, 40 + ci, mut)
        assert all(crc for _, crc, _ in res), (cmb, [crc for _, crc, _ in res])
        for g, _, _ in res:
            L = ltelib.uci_layout(g)
            assert L.G == (144 * g.L_prb - L.Qp_cqi - L.Qp_ri) * g.qm
            assert (L.Qp_ack > 0) == (g.nof_ack > 0) and (L.Qp_ri > 0) == (g.ri_len > 0) and (L.Qp_cqi > 0) == (g.cqi_len > 0)
    # a receiver that ignores the CQI (wrong G, shifted codeword) must fail: the multiplexing really moves the data
    o = ltelib.Oracle(cell)
    s = ltelib.Sim(cell=cell, seed=77, snr_db=34.0, nof_ues=1)
    rng = np.random.default_rng(77)
    gr = ltelib.make_ul_grants(cell, rng, 2)
    for g in gr:
        g.cqi_len, g.I_offset_cqi = 30, 8
    x, pl, off = ltelib.sim_ul_subframe(s, 3, ucfg, gr)
    for g in gr:
        g.cqi_len = 0
    _, ref = ltelib.oracle_ul(o, ucfg, 3, gr, x)
    assert not any(crc for _, _, crc, _, _ in ref)


def test_pusch_group_and_sequence_hopping_roundtrip(infra):
    """DMRS base-sequence group hopping and sequence hopping (36.211 5.5.1.3 / 5.5.1.4; dmrs_cfg.group_hopping_en / sequence_hopping_en from SIB2,
    src/src/ULSchedule.cc:143-146): transmitter and receiver agree on u, v per slot; a receiver with the wrong setting loses the channel estimate"""
    from ltelib import UlCfg
    cell = Cell(50, 1, 301, 1)
    for gh, sh in ((1, 0), (0, 1), (1, 1)):
        ucfg = UlCfg(n_dmrs1=2, delta_ss=11, group_hopping=gh, seq_hopping=sh)
        res = _ul_roundtrip(cell, ucfg, 50 + 2 * gh + sh, lambda g, rng: None, min_prb=6)
        assert all(crc for _, crc, _ in res) and len(res) >= 6
    S = ltelib.sim()
    S.lte_pusch_uv.argtypes = [C.POINTER(Cell), C.POINTER(UlCfg), C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    u, v = C.c_uint32(), C.c_uint32()
    us, vs = set(), set()
    for ns in range(20):
        S.lte_pusch_uv(C.byref(cell), C.byref(UlCfg(delta_ss=11, group_hopping=1)), ns, 72, C.byref(u), C.byref(v))
        us.add(u.value)
        assert v.value == 0
        S.lte_pusch_uv(C.byref(cell), C.byref(UlCfg(delta_ss=11, seq_hopping=1)), ns, 72, C.byref(u), C.byref(v))
        vs.add(v.value)
        assert u.value == (301 % 30 + 11) % 30
        S.lte_pusch_uv(C.byref(cell), C.byref(UlCfg(delta_ss=11, seq_hopping=1)), ns, 60, C.byref(u), C.byref(v))
        assert v.value == 0            # below 6 PRB there is one base sequence per group
    assert len(us) > 5 and vs == {0, 1}
    # mismatch: hopping transmitter, static receiver
    o = ltelib.Oracle(cell)
    s = ltelib.Sim(cell=cell, seed=9, snr_db=34.0, nof_ues=1)
    gr = ltelib.make_ul_grants(cell, np.random.default_rng(9), 2, min_prb=6)
    x, pl, off = ltelib.sim_ul_subframe(s, 3, UlCfg(n_dmrs1=2, delta_ss=11, group_hopping=1), gr)
    _, ref = ltelib.oracle_ul(o, UlCfg(n_dmrs1=2, delta_ss=11), 3, gr, x)
    assert not all(crc for _, _, crc, _, _ in ref)


def test_pusch_type1_hopping_and_timing_offset(infra):
    """type-1 PUSCH hopping (slot 1 on other PRBs, 36.213 8.4.1 as restated by ul_sniffer_ra_ul_grant_to_grant_prb_allocation,
    lib/src/phy/falcon_phch/ul_sniffer_pusch.c:48-80) and the timing-offset estimate from the DMRS phase slope (chest_res.ta_us, UL_Sniffer_PUSCH.cc:424,574)"""
    from ltelib import UlCfg
    cell = Cell(50, 1, 5, 1)
    ucfg = UlCfg(n_dmrs1=1, delta_ss=0, n_rb_ho=4)
    o = ltelib.Oracle(cell)
    s = ltelib.Sim(cell=cell, seed=12, snr_db=32.0, nof_ues=1)
    S = ltelib.sim()
    S.lte_ul_dci_to_grant_hop.argtypes = [C.POINTER(Cell), C.POINTER(UlCfg), C.POINTER(ltelib.Dci), C.c_int, C.POINTER(ltelib.UlGrant)]
    N, rivb = 50, 11
    seen = set()
    for hb, (L, S0) in zip((0, 1, 2, 0, 2), ((6, 4), (5, 30), (8, 10), (10, 20), (4, 2))):
        d = ltelib.Dci()
        d.format, d.rnti, d.alloc_type, d.hop = 0, 0x4000 + hb, 2, 1
        riv = N * (L - 1) + S0
        assert riv < (1 << (rivb - 2))
        d.riv = (hb << (rivb - 2)) | riv
        d.mcs[0], d.n_dmrs = 12, 3
        g = ltelib.UlGrant()
        assert S.lte_ul_dci_to_grant_hop(C.byref(cell), C.byref(ucfg), C.byref(d), 1, C.byref(g)) == 0
        nrb = N - 4
        want = {0: (nrb // 4 + S0) % nrb, 1: (nrb + S0 - nrb // 4) if S0 < nrb // 4 else S0 - nrb // 4, 2: (nrb // 2 + S0) % nrb}[hb]
        assert (g.hop, g.n_prb, g.n_prb_slot1, g.L_prb) == (1, S0, want, L)
        seen.add(hb)
        g.ta_us = 0.4 + 0.3 * hb
        x, pl, off = ltelib.sim_ul_subframe(s, 6, ucfg, [g])
        _, ref = ltelib.oracle_ul(o, ucfg, 6, [g], x)
        r, opl, crc, ch, _ = ref[0]
        assert r == 0 and crc and np.array_equal(opl[:g.tbs // 8], pl[:g.tbs // 8])
        assert abs(ch.ta_us - g.ta_us) < 0.05, (ch.ta_us, g.ta_us)
    assert seen == {0, 1, 2}
    # no offset -> estimate near zero; negative offset (early UE) keeps its sign
    for ta in (0.0, -0.7):
        gr = ltelib.make_ul_grants(cell, np.random.default_rng(4), 2)
        for g in gr:
            g.ta_us = ta
        x, pl, off = ltelib.sim_ul_subframe(s, 7, ucfg, gr)
        _, ref = ltelib.oracle_ul(o, ucfg, 7, gr, x)
        for r, opl, crc, ch, _ in ref:
            assert crc and abs(ch.ta_us - ta) < 0.05


@pytest.mark.parametrize("mcs,snr_ok,snr_fail", [(9, 7.0, 0.0), (16, 14.0, 6.0), (28, 25.0, 17.0)])
def test_receiver_sensitivity_is_in_the_expected_range(infra, mcs, snr_ok, snr_fail):
    """A yardstick from outside for the whole oracle receiver (estimator, equaliser, soft demodulator, rate-dematching, turbo): on a flat single-antenna channel
    the 10 % block-error points of LTE link-level tables are near 2.5 dB (QPSK, rate 0.58), 8 dB (16QAM, 0.56) and 18.5 dB (64QAM, 0.89); this receiver, which
    estimates the channel from the CRS as srsRAN does, was measured about 1 dB later (16QAM: 10 % at 8.8 dB; 7.8 dB when it is given the channel, i.e. the
    demodulator and decoder sit on the theoretical curve and the rest is estimation noise).  Every block must decode some dB above that and none
    some dB below the ideal point: a wrong LLR scale, constellation threshold or noise estimate costs far more than the margin left here."""
    from helpers import make_capture, truth_grants, oracle_frontend
    cell = Cell(25, 1, 7, 1)
    res = {}
    for snr in (snr_ok, snr_fail):
        sim, iq, tti, truths, payloads = make_capture(cell, 6, seed=100 + mcs, cfi=2, nof_ues=4, dl_min=2, dl_max=2, tm=1, mcs_min=mcs, mcs_max=mcs, snr_db=snr, full_band=1)
        o = Oracle(cell)
        fe = oracle_frontend(o, iq, tti)
        ok = n = 0
        for sf, d, g in truth_grants(cell, truths, tti):
            r, pl, okk = o.pdsch_decode(int(tti[sf]) % 10, fe[sf]["cfi"], d.rnti, g, fe[sf]["sym"], fe[sf]["ce"], 8)
            n += 1
            ok += int(okk[0])
        res[snr] = (ok, n)
    assert res[snr_ok][0] == res[snr_ok][1] >= 10 and res[snr_fail][0] == 0, res


def test_channel_estimate_on_noiseless_channels(infra):
    """the CRS estimator of the oracle on channels it should get exactly: without noise a flat path of gain g gives |ce| = g sqrt(N) on every RE (direct paths 1.0,
    cross paths 0.35 in the simulator), constant over the 14 symbols; a path delayed by d samples gives a phase ramp of -2 pi d / N per sub-carrier.  Pins pilot
    positions and values, the interpolation over frequency and time and the port / antenna ordering against physics rather than against the transmitter's code."""
    from helpers import make_capture, oracle_frontend
    cell = Cell(50, 2, 21, 2)
    N = cell.fft()
    for delay in (0, 3):
        sim, iq, tti, truths, payloads = make_capture(cell, 2, seed=7, cfi=2, nof_ues=4, dl_min=2, dl_max=3, tm=2, mcs_min=4, mcs_max=10, snr_db=100.0, chan_delay=delay)
        fe = oracle_frontend(Oracle(cell), iq, tti)
        ce = fe[1]["ce"].reshape(2, 2, 14, 12 * cell.nof_prb)                # [port][antenna][symbol][sub-carrier]
        delays = []
        for p in range(2):
            for a in range(2):
                c = ce[p, a]
                gain = 1.0 if a == p else 0.35
                assert abs(np.abs(c).mean() / (gain * np.sqrt(N)) - 1) < (2e-2 if delay else 5e-3), (delay, p, a, np.abs(c).mean())   # linear interpolation between
                # pilots 6 sub-carriers apart shortens a rotating phasor a little
                step = np.angle(c[:, 1:] / c[:, :-1])
                assert step.std() < 2e-3                                     # one slope over the whole band and every symbol
                d = -step.mean() * N / (2 * np.pi)
                assert abs(d - round(d)) < 0.05 and 0 <= round(d) <= delay, (delay, p, a, d)
                delays.append(int(round(d)))
                if delay == 0:
                    assert np.abs(c - c.mean()).max() < 2e-3 * np.abs(c).mean()
        if delay:
            assert max(delays) >= 1                                          # at least one path was really delayed
