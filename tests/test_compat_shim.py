"""Tier-2 shim (include/ltephy_srsran_compat.h, pytest -m gpu): examples/compat_check.cpp calls the srsRAN / FALCON names the way
the reference does, one subframe at a time; everything it gets back must equal what the batched tier-1 path returns for the same
capture: CFI, SNR, sf_symbols / ce / PDCCH LLRs, the result of srsran_pdcch_dci_decode for every (location, size), and the
transport blocks of srsran_ue_dl_decode_pdsch."""
import os
import struct
import subprocess
import numpy as np
import pytest
from ltelib import Cell
from helpers import make_capture, to_phy_grant, truth_grants
from ltesniffer_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "compat_check")


def fnv(b):
    h = 1469598103934665603
    for x in np.frombuffer(b, np.uint8).tolist():
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_shim_equals_tier1(infra, phylib, tmp_path):
    capi.load_library()
    subprocess.run(["make", "-s", "-C", ROOT, "examples/compat_check"], check=True)
    cell = Cell(25, 2, 150, 2)
    n = 4
    sim, iq, tti, truths, payloads = make_capture(cell, n, seed=14, cfi=2, nof_ues=5, dl_min=1, dl_max=3, tm=13, mcs_min=2, mcs_max=12, snr_db=25.0, tti0=9)
    f = tmp_path / "cap.cf32"
    iq.astype(np.complex64).tofile(str(f))
    tg = truth_grants(cell, truths, tti)
    grants = [to_phy_grant(sf, d.rnti, g) for sf, d, g in tg]
    with open(tmp_path / "grants.bin", "wb") as gf:
        for (sf, d, g), pg in zip(tg, grants):
            gf.write(struct.pack("<II", sf, d.rnti) + bytes(pg))
    # tier 1
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=n, turbo_max_iter=8)
    phy.submit_iq(iq, tti)
    info, cands = phy.get_phase_a()
    g = 14 * 12 * cell.nof_prb
    sym = phy.tap(capi.TAP_SYM, (n, cell.nof_rx, g), np.complex64)
    ce = phy.tap(capi.TAP_CE, (n, cell.nof_ports, cell.nof_rx, g), np.complex64)
    llr = phy.tap(capi.TAP_LLR, (n, capi.LLR_STRIDE), np.float32)
    phy.submit_grants(grants)
    res, pl = phy.get_phase_b()
    sizes, sidx = phy.sizes()
    distinct = []
    for s in sizes:
        if s not in distinct:
            distinct.append(s)
    out = tmp_path / "out.bin"
    r = subprocess.run([EXE, str(f), str(cell.nof_prb), str(cell.nof_ports), str(cell.cell_id), str(cell.nof_rx), str(n), str(int(tti[0])),
                        str(tmp_path / "grants.bin"), str(out)] + [str(s) for s in distinct], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    b = out.read_bytes()
    off = 0
    ntb = 0
    for i in range(n):
        cfi, snr, noise, cfo, hs, hc, ncce = struct.unpack_from("<IfffQQI", b, off)
        off += struct.calcsize("<IfffQQI")
        assert cfi == info[i].cfi and np.float32(snr) == np.float32(info[i].snr_db) and np.float32(noise) == np.float32(info[i].noise_avg)
        assert np.float32(cfo) == np.float32(info[i].cfo)
        es = 0
        for a in range(cell.nof_rx):
            es ^= (fnv(sym[i, a].tobytes()) * (a + 1)) & 0xFFFFFFFFFFFFFFFF
        ec = 0
        for p in range(cell.nof_ports):
            for a in range(cell.nof_rx):
                ec ^= (fnv(ce[i, p, a].tobytes()) * (p * 2 + a + 1)) & 0xFFFFFFFFFFFFFFFF
        assert hs == es and hc == ec, "sf_symbols / ce handed out by the shim differ from the device buffers"
        assert ncce == phy.nof_cce(cfi)
        got_llr = np.frombuffer(b, np.float32, 72 * ncce, off)
        off += 4 * 72 * ncce
        assert np.array_equal(got_llr.view(np.uint32), llr[i, :72 * ncce].view(np.uint32))
        nc, Ls = phy.locations(cfi)
        for li in range(len(nc)):
            for nb in distinct:
                crc, bits = struct.unpack_from("<HQ", b, off)
                off += 10
                c = cands[i, li, distinct.index(nb)]
                assert (crc, bits) == ((int(c["rnti"]), int(c["bits"])) if c["valid"] else (0, 0)), (i, li, nb)
        for gi, (sf, d, gg) in enumerate(tg):
            if sf != i:
                continue
            for t in range(2):
                crc, nby = struct.unpack_from("<BI", b, off)
                off += 5
                data = b[off:off + nby]
                off += nby
                rr = res[2 * gi + t]
                assert nby == rr.payload_len and bool(crc) == bool(rr.crc)
                assert data == bytes(pl[rr.payload_off:rr.payload_off + nby])
                ntb += 1 if nby else 0
    assert off == len(b) and ntb >= n
    phy.close()
