"""examples/offline_decode.cpp: the reference's offline mode written in C++ against the C-ABI (IQ file -> MAC-LTE pcap + DCI trace).
CPU: it builds and links against libltephy_b200.so, and without a GPU it fails loudly (no fallback).  GPU: on a synthetic capture its
pcap carries exactly the transport blocks the Python-driven pipeline decodes, which are the transmitter's."""
import os
import struct
import subprocess
import numpy as np
import pytest
from ltelib import Cell
from helpers import make_capture
from ltesniffer_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "offline_decode")


def _build():
    capi.load_library()
    subprocess.run(["make", "-s", "-C", ROOT, "examples/offline_decode"], check=True)
    assert os.path.exists(EXE)


def test_example_builds_and_refuses_to_run_without_a_gpu(infra, tmp_path):
    _build()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present: covered by the gpu test")
    iq = tmp_path / "x.cf32"
    iq.write_bytes(b"\0" * 8 * 7680)
    r = subprocess.run([EXE, str(iq), "25", "1", "1", "1", str(tmp_path / "o.pcap"), str(tmp_path / "o.tsv")], capture_output=True, text=True)
    assert r.returncode == 1 and "ltephy_create" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_example_matches_python_pipeline(infra, phylib, tmp_path):
    _build()
    cell = Cell(50, 2, 301, 2)
    n = 40
    sim, iq, tti, truths, payloads = make_capture(cell, n, seed=3, cfi=3, nof_ues=10, dl_min=3, dl_max=5, ul_min=1, ul_max=2, tm=13, mcs_min=2, mcs_max=18, snr_db=25.0)
    f = tmp_path / "cap.cf32"
    iq.astype(np.complex64).tofile(str(f))
    pc, tsv = str(tmp_path / "o.pcap"), str(tmp_path / "o.tsv")
    r = subprocess.run([EXE, str(f), "50", "2", "301", "2", pc, tsv, "16"], capture_output=True, text=True)     # batch of 16: three calls
    assert r.returncode == 0, r.stderr
    # the same capture through the Python-driven pipeline in the same batches (the RNTI history carries over between calls)
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=16, turbo_max_iter=8, flags=capi.FLAG_SKIP_LOW_POWER)
    srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
    expect, ndci = [], 0
    for lo in range(0, n, 16):
        info, dcis, tbs, payload = capi.decode_subframes(phy, srch, iq[lo:lo + 16], tti[lo:lo + 16])
        ndci += len(dcis)
        for i, d in enumerate(dcis):
            for t in range(2):
                rr = tbs[2 * i + t]
                if rr.crc and rr.payload_len:
                    expect.append((int(d["rnti"]), int(tti[lo + int(d["sf"])]), bytes(payload[rr.payload_off:rr.payload_off + rr.payload_len])))
    phy.close()
    b = open(pc, "rb").read()
    off, got = 24, []
    while off < len(b):
        ts, tu, il, ol = struct.unpack("<IIII", b[off:off + 16])
        rec = b[off + 16:off + 16 + il]
        off += 16 + il
        sfn_sf = (rec[10] << 8) | rec[11]
        got.append(((rec[4] << 8) | rec[5], (sfn_sf >> 4) * 10 + (sfn_sf & 15), rec[19:]))
    assert got == expect and len(got) > n
    lines = open(tsv).read().splitlines()
    assert len(lines) >= ndci * 0.9 and all(len(l.split("\t")) == 20 for l in lines)
    assert ("subframes %d " % n) in r.stdout


def test_ul_mode_example_builds_and_refuses_to_run_without_a_gpu(infra, tmp_path):
    """examples/offline_ul.cpp: the UL mode as a C++ caller of the C-ABI (ltephy_search_set_ul_mode, ltephy_rar_unpack, ltephy_ul_grants_from_dcis,
    ltephy_submit_ul ...): it compiles against the headers as they are, links, and without a GPU fails in ltephy_create"""
    capi.load_library()
    subprocess.run(["make", "-s", "-C", ROOT, "examples/offline_ul"], check=True)
    exe = os.path.join(ROOT, "examples", "offline_ul")
    assert os.path.exists(exe)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    for name in ("d.cf32", "u.cf32"):
        (tmp_path / name).write_bytes(b"\0" * 8 * 7680)
    r = subprocess.run([exe, str(tmp_path / "d.cf32"), str(tmp_path / "u.cf32"), "25", "1", "1", str(tmp_path / "o.pcap")], capture_output=True, text=True)
    assert r.returncode == 1 and "ltephy_create" in r.stderr, (r.returncode, r.stderr)
    assert subprocess.run([exe], capture_output=True).returncode == 2
