"""GPU parity of the PUSCH path (pytest -m gpu): UL OFDM demodulation and every PUSCH grant's channel estimate
figures, transport-block bytes and CRC against the CPU oracle, bit-exact; ground truth from the synthetic UEs."""
import numpy as np
import pytest
import ltelib
from ltelib import Cell, Sim, Oracle, UlCfg
from helpers import feq, describe_mismatch
from ltesniffer_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cell,snr,nsf,ngr,table", [(Cell(50, 1, 17, 1), 25.0, 3, 4, 1), (Cell(100, 2, 301, 2), 23.0, 3, 6, 1), (Cell(25, 1, 5, 1), 20.0, 2, 2, 1),
                                                      (Cell(50, 1, 9, 1), 34.0, 3, 4, 2)])     # table 2: 36.213 Table 8.6.1-3, Qm up to 8
def test_pusch_bit_exact(infra, phylib, cell, snr, nsf, ngr, table):
    s = Sim(cell=cell, seed=21, snr_db=snr, nof_ues=1, chan_delay=3)
    o = Oracle(cell)
    rng = np.random.default_rng(cell.nof_prb)
    ucfg = UlCfg(n_dmrs1=3, delta_ss=2)
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=nsf)
    phy.set_ul_cfg(3, 2)
    iq = np.zeros((nsf, s.sf_len), np.complex64)
    tti = np.arange(4, 4 + nsf, dtype=np.uint32)
    grants_o, pls, offs, grants_p = [], [], [], []
    for i in range(nsf):
        gr = ltelib.make_ul_grants(cell, rng, ngr, table=table)
        x, pl, off = ltelib.sim_ul_subframe(s, int(tti[i]), ucfg, gr)
        iq[i] = x
        grants_o.append(gr), pls.append(pl), offs.append(off)
        for g in gr:
            grants_p.append(capi.UlGrant(sf=i, rnti=g.rnti, qm=g.qm, rv=g.rv, L_prb=g.L_prb, n_prb=g.n_prb, n_dmrs2=g.n_dmrs2, tbs=g.tbs))
    res, ch, payload = phy.decode_ul(iq, tti, grants_p)
    ulsym = phy.tap(capi.TAP_UL_SYM, (nsf, 14 * phy.nsc), np.complex64)
    k = 0
    nok = 0
    for i in range(nsf):
        sym, ref = ltelib.oracle_ul(o, ucfg, int(tti[i]), grants_o[i], iq[i])
        assert feq(ulsym[i], sym), describe_mismatch(ulsym[i], sym, "UL grid sf %d" % i)
        for g, (r, opl, ocrc, och, _), off in zip(grants_o[i], ref, offs[i]):
            assert r == 0
            rr = res[k]
            nby = g.tbs // 8
            assert (ch[k].noise, ch[k].rsrp, ch[k].snr_db) == (och.noise, och.rsrp, och.snr_db), (k, ch[k].noise, och.noise)
            assert rr.crc == ocrc and rr.payload_len == nby
            assert np.array_equal(payload[rr.payload_off:rr.payload_off + nby], opl[:nby])
            if rr.crc:
                assert np.array_equal(payload[rr.payload_off:rr.payload_off + nby], pls[i][off:off + nby])
                nok += 1
            k += 1
    assert nok >= k - 1
    phy.close()
