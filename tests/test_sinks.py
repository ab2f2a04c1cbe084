"""Result back-end wire formats (include/ltephy_sinks.h) against the reference's OWN output files -- the one place where the
reference ships golden bytes: pcap_file_example/*.pcap are captures written by LTESniffer_pcap_writer.  Every record of those
files is parsed (context + PDU + timestamp), written again through ltephy_pcap_write, and the resulting file must be
byte-identical (run where /root/reference is mounted; elsewhere the two records quoted below as known answers still pin the layout).
Also: the DCI trace line against a restatement of DCIToFile::printDCICollection's format string."""
import ctypes as C
import os
import struct
import numpy as np
import pytest
import ltelib
from ltelib import Cell
from ltesniffer_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/pcap_file_example"
FILES = [os.path.join(REF_DIR, f) for f in ("ltesniffer_dl_mode.pcap", "ltesniffer_ul_mode.pcap", "api_collector.pcap") if os.path.exists(os.path.join(REF_DIR, f))]
# the first two records of the reference's pcap_file_example/api_collector.pcap (record header + MAC-LTE context + PDU), as known answers
KAT = ["bece44643c9e02001a0000001a000000010003020048030000043a4207010a000f0001005bd3064519c6",
       "bece44647ab502003300000033000000010103020048030000043a4807010a000f00013c20141f5bd3064519c660129b2e661e82f2e0ccc860d30000990a0003e00000"]


def _lib():
    L = capi.load_library()
    P = C.c_void_p
    L.ltephy_pcap_open.argtypes = [C.c_char_p]
    L.ltephy_pcap_open.restype = P
    L.ltephy_pcap_close.argtypes = [P]
    L.ltephy_pcap_write.argtypes = [P, P, C.c_uint32, C.c_uint16, C.c_uint8, C.c_uint8, C.c_uint32, C.c_int, C.c_uint16, C.c_uint32, C.c_uint32]
    L.ltephy_rnti_type.argtypes = [C.c_uint16]
    L.ltephy_rnti_type.restype = C.c_uint8
    L.ltephy_pcap_write_dl_batch.argtypes = [P, P, P, C.c_uint32, P, P, C.c_uint16, C.c_uint32, C.c_uint32]
    L.ltephy_pcap_write_ul_batch.argtypes = [P, P, P, C.c_uint32, P, P, C.c_uint16, C.c_uint32, C.c_uint32]
    L.ltephy_dci_trace_line.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_char_p, C.c_size_t]
    return L


def parse(path):
    b = open(path, "rb").read()
    assert struct.unpack("<IHHiIII", b[:24]) == (0xa1b2c3d4, 2, 4, 0, 0, 65535, 147)
    off, recs = 24, []
    while off < len(b):
        ts, tu, il, ol = struct.unpack("<IIII", b[off:off + 16])
        off += 16
        r = b[off:off + il]
        off += il
        assert il == ol and r[0] == 1 and r[3] == 2 and r[6] == 3 and r[9] == 4 and r[12] == 7 and r[14] == 0x0a and r[15] == 0 and r[16] == 0x0f \
            and r[17] == 0 and r[18] == 1, "unexpected MAC-LTE context layout"
        sfn_sf = (r[10] << 8) | r[11]
        recs.append(dict(ts=ts, tu=tu, direction=r[1], rnti_type=r[2], rnti=(r[4] << 8) | r[5], ueid=(r[7] << 8) | r[8], tti=(sfn_sf >> 4) * 10 + (sfn_sf & 15),
                         crc=r[13], pdu=r[19:]))
    return b, recs


def test_pcap_known_answer_records(infra, tmp_path):
    L = _lib()
    out = str(tmp_path / "kat.pcap")
    p = L.ltephy_pcap_open(out.encode())
    for h in KAT:
        r = bytes.fromhex(h)
        ts, tu, il, ol = struct.unpack("<IIII", r[:16])
        c, pdu = r[16:35], np.frombuffer(r[35:], np.uint8)
        sfn_sf = (c[10] << 8) | c[11]
        assert L.ltephy_pcap_write(p, pdu.ctypes.data_as(C.c_void_p), len(pdu), (c[4] << 8) | c[5], c[2], c[1], (sfn_sf >> 4) * 10 + (sfn_sf & 15), c[13],
                                   (c[7] << 8) | c[8], ts, tu) == 0
    L.ltephy_pcap_close(p)
    b = open(out, "rb").read()
    assert b[:24].hex() == "d4c3b2a1020004000000000000000000ffff000093000000" and b[24:].hex() == "".join(KAT)


@pytest.mark.skipif(not FILES, reason="the reference's example captures are not mounted")
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(p) for p in FILES])
def test_pcap_records_reproduce_the_references_files(infra, tmp_path, path):
    L = _lib()
    orig, recs = parse(path)
    assert len(recs) >= 20
    out = str(tmp_path / "out.pcap")
    p = L.ltephy_pcap_open(out.encode())
    assert p
    for r in recs:
        buf = np.frombuffer(r["pdu"], np.uint8)
        assert L.ltephy_pcap_write(p, buf.ctypes.data_as(C.c_void_p), len(buf), r["rnti"], r["rnti_type"], r["direction"], r["tti"], r["crc"], r["ueid"],
                                   r["ts"], r["tu"]) == 0
        if r["direction"] == 1:      # downlink records: the RNTI class the reference chose is the one ltephy_rnti_type gives
            assert L.ltephy_rnti_type(r["rnti"]) == r["rnti_type"], hex(r["rnti"])
    L.ltephy_pcap_close(p)
    assert open(out, "rb").read() == orig


def test_dl_batch_writes_crc_passing_blocks_only(infra, tmp_path):
    L = _lib()
    dcis = np.zeros(3, capi.DCI_DTYPE)
    dcis["sf"], dcis["rnti"] = [0, 1, 1], [0xFFFF, 0x1234, 0x0005]
    tbs = (capi.TbResult * 6)()
    payload = np.arange(64, dtype=np.uint8)
    for i, (crc, off, ln) in enumerate([(1, 0, 10), (0, 0, 0), (0, 10, 8), (2, 18, 6), (1, 24, 4), (0, 0, 0)]):
        tbs[i].crc, tbs[i].payload_off, tbs[i].payload_len = crc, off, ln
    tti = np.array([4301, 4302], np.uint32)
    out = str(tmp_path / "b.pcap")
    p = L.ltephy_pcap_open(out.encode())
    n = L.ltephy_pcap_write_dl_batch(p, tti.ctypes.data_as(C.c_void_p), dcis.ctypes.data_as(C.c_void_p), 3, tbs, payload.ctypes.data_as(C.c_void_p), 7, 100, 5)
    L.ltephy_pcap_close(p)
    assert n == 3
    _, recs = parse(out)
    assert [(r["rnti"], r["rnti_type"], r["tti"], len(r["pdu"]), r["ueid"], r["direction"]) for r in recs] == \
           [(0xFFFF, 4, 4301, 10, 7, 1), (0x1234, 3, 4302, 6, 7, 1), (0x0005, 2, 4302, 4, 7, 1)]
    assert recs[1]["pdu"] == bytes(range(18, 24))


def test_ul_batch_writes_crc_passing_blocks_only(infra, tmp_path):
    L = _lib()
    grants = (capi.UlGrant * 3)(capi.UlGrant(sf=0, rnti=0x46, qm=2, L_prb=3, tbs=208), capi.UlGrant(sf=1, rnti=0x47, qm=4, L_prb=4, tbs=256),
                                capi.UlGrant(sf=1, rnti=0x48, qm=2, L_prb=3, tbs=208))
    res = (capi.TbResult * 3)()
    payload = np.arange(200, dtype=np.uint8)
    for i, (crc, off, ln) in enumerate([(1, 0, 26), (0, 26, 32), (1, 58, 26)]):
        res[i].crc, res[i].payload_off, res[i].payload_len = crc, off, ln
    tti = np.array([1234, 1235], np.uint32)
    out = str(tmp_path / "u.pcap")
    p = L.ltephy_pcap_open(out.encode())
    assert L.ltephy_pcap_write_ul_batch(p, tti.ctypes.data_as(C.c_void_p), grants, 3, res, payload.ctypes.data_as(C.c_void_p), 0, 9, 8) == 2
    L.ltephy_pcap_close(p)
    _, recs = parse(out)
    assert [(r["rnti"], r["rnti_type"], r["direction"], r["tti"], len(r["pdu"])) for r in recs] == [(0x46, 3, 0, 1234, 26), (0x48, 3, 0, 1235, 26)]


def test_dci_trace_line_format(infra):
    """product line == the reference's format strings (SubframeInfoConsumer.cc:86-95,104-113,125-133) filled from the oracle's unpack"""
    L = _lib()
    capi._bind_search(L)
    S = infra.sim()
    S.lte_ul_dci_to_grant.argtypes = [C.POINTER(Cell), C.POINTER(ltelib.Dci), C.c_int, C.POINTER(ltelib.UlGrant)]
    rng = np.random.default_rng(3)
    cell = Cell(100, 2, 3, 2)
    srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
    out = C.create_string_buffer(512)
    nlines = 0
    for _ in range(600):
        f = int(rng.choice([0, 1, 2, 7]))
        nb = S.lte_dci_sizeof(C.byref(cell), f)
        bits = rng.integers(0, 2, nb).astype(np.uint8)
        if f == 0:
            bits[0], bits[1] = 0, 0
        if f == 2:
            bits[0] = 1
        rnti, tti, cfi = int(rng.integers(11, 0xFFF3)), int(rng.integers(0, 10240)), int(rng.integers(1, 4))
        v = 0
        for i, b in enumerate(bits):
            v |= int(b) << (63 - i)
        row = np.zeros(1, capi.DCI_DTYPE)
        row["rnti"], row["format"], row["nof_bits"], row["bits"], row["ncce"], row["L"], row["histogram_value"] = rnti, f, nb, v, 8, 2, 17
        n = L.ltephy_dci_trace_line(srch.h, row.ctypes.data_as(C.c_void_p), tti, cfi, 0, 1700000000, 42, out, 512)
        hexs = "".join("%02x" % ((v >> (56 - 8 * i)) & 0xFF) for i in range((nb + 7) // 8))
        d = ltelib.Dci()
        r0 = S.lte_dci_unpack(C.byref(cell), f, rnti, ltelib.ptr(bits), nb, C.byref(d))
        if f == 0:
            g = ltelib.UlGrant()
            ok = r0 == 0 and S.lte_ul_dci_to_grant(C.byref(cell), C.byref(d), 1, C.byref(g)) in (0, -2)
            if not ok or g.L_prb == 0:
                continue
            exp = "%d.%06d\t%04d\t%d\t%d\t0\t%d\t%d\t%d\t%d\t%d\t0\t%d\t-1\t%d\t%d\t%d\t%d\t%d\t%d\t%s\n" % (
                1700000000, 42, tti // 10, tti % 10, rnti, d.mcs[0], g.L_prb, g.tbs if d.mcs[0] < 29 else 0, -1, -1, d.ndi[0], tti % 8, 8, 2, cfi, 17, nb, hexs)
        else:
            r0, d, g = ltelib.unpack_and_grant(cell, f, rnti, bits, tti % 10, cfi, 0)
            if r0 != 0:
                assert n < 0
                continue
            t0 = max(0, g.tb[0].tbs)
            t1 = t0       # the reference prints the first block's size in both columns (convert_dl_grant, falcon_dci.c:608-612)
            two = f >= 6
            exp = "%d.%06d\t%04d\t%d\t%d\t1\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%s\n" % (
                1700000000, 42, tti // 10, tti % 10, rnti, d.mcs[0], g.nof_prb, t0 + t1 if two else t0, t0 if two else -1, t1 if two else -1, f + 1,
                d.ndi[0], d.ndi[1] if two else -1, d.pid, 8, 2, cfi, 17, nb, hexs)
        assert n == len(exp) and out.value.decode() == exp, (f, out.value.decode(), exp)
        nlines += 1
    assert nlines > 200
