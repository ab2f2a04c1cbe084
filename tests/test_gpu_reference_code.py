"""pytest -m gpu: the reference's OWN unmodified blind search (src/src/DCISearch.cc + lib/src/phy/falcon_phch/*.c, compiled against
compat/srsran into oracle/_ref/libfalcon_ref.so) walking the candidate tables the GPU produced, against ltephy_search_batch and the
survivor-form walk over the same GPU output; and the srsRAN-named tier-2 objects (srsran_ue_dl_t ...) used by that search reading
per-CCE power and LLRs of the GPU."""
import numpy as np
import pytest
from ltelib import Cell
from helpers import make_capture
from ltesniffer_b200 import capi
from test_reference_code import RefWalk, compare_grants, needs_ref

pytestmark = pytest.mark.gpu


@needs_ref
@pytest.mark.parametrize("cell,n,kw", [
    (Cell(50, 2, 301, 2), 40, dict(seed=3, cfi=3, nof_ues=10, dl_min=3, dl_max=5, ul_min=1, ul_max=2, tm=13, mcs_min=2, mcs_max=18, snr_db=25.0, chan_delay=5)),
    (Cell(100, 2, 7, 2), 30, dict(seed=2, cfi=3, nof_ues=150, dl_min=8, dl_max=12, tm=3, mcs_min=17, mcs_max=26, snr_db=28.0, full_band=1)),
])
def test_reference_walk_over_gpu_tables(infra, phylib, cell, n, kw):
    sim, iq, tti, truths, payloads = make_capture(cell, n, **kw)
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=n, turbo_max_iter=8)   # every location decoded
    phy.submit_iq(iq, tti)
    info, cands = phy.get_phase_a()
    info_c, comp = phy.get_phase_a_compact()
    llr = phy.tap(capi.TAP_LLR, (n, capi.LLR_STRIDE), np.float32)
    ref = RefWalk(cell)
    srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
    srch_c = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
    ref.L.refwalk_config(ref.h, 1, 0, 10)
    srch.config(1, 0, 10)
    srch_c.config(1, 0, 10)
    total = 0
    for i in range(n):
        want = ref.subframe(info[i], cands[i], llr[i])
        got = srch.subframe(info[i], cands[i])
        got_c = srch_c.subframe_compact(info[i], comp[i:i + 1])
        assert got_c is not None and len(got_c) == len(got) and all(np.array_equal(got_c[k], got[k]) for k in got.dtype.names)
        for is_ul in (False, True):
            a = [d for d in got if (d["format"] == 0) == is_ul]
            b = [r for r in want if (r.format == 0) == is_ul]
            assert len(a) == len(b), (i, is_ul, len(a), len(b))
            for d, r in zip(a, b):
                assert (int(d["rnti"]), int(d["format"]), int(d["L"]), int(d["ncce"]), int(d["nof_bits"]), int(d["histogram_value"])) == \
                       (r.rnti, r.format, r.L, r.ncce, r.nof_bits, r.histval), (i, is_ul)
                assert np.array_equal(capi.cand_bits(d["bits"], r.nof_bits), np.frombuffer(bytes(r.bits), np.uint8)[:r.nof_bits])
                compare_grants(srch, cell, info[i], d, r)
        total += len(got)
    rs, ps = ref.stats(), srch.stats()
    assert (rs.nof_decoded_locations, rs.nof_cce, rs.nof_missed_cce, rs.nof_subframes, rs.nof_locations) == \
           (ps.nof_decoded_locations, ps.nof_cce, ps.nof_missed_cce, ps.nof_subframes, ps.nof_locations)
    assert total >= n
    ref.close()
    phy.close()
