"""3GPP table checks (CPU): structural validation of include/lte_tables.h plus the anchors that exist in
the reference tree (row 32A, format-1C TBS table)."""
import os
import re
import sys
import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import check_tables  # noqa: E402

REF = "/root/reference"


def test_tables_structurally_valid():
    assert check_tables.check() == []


def test_known_answers():
    f1, f2, tbs, f1c = check_tables.load()
    T = np.array(tbs).reshape(34, 110)
    assert T[26, 99] == 75376 and T[33, 99] == 97896 and T[0, 0] == 16 and T[26, 0] == 712
    assert check_tables.segm(75376)[:2] == (13, 5824) and check_tables.segm(97896)[:2] == (16, 6144) and check_tables.segm(36696)[:2] == (6, 6144)
    assert (f1[0], f2[0]) == (3, 10) and (f1[187], f2[187]) == (263, 480)                      # K = 40, K = 6144


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_anchors_in_reference_tree():
    _, _, tbs, f1c = check_tables.load()
    T = np.array(tbs).reshape(34, 110)
    src = open(os.path.join(REF, "lib/src/phy/falcon_phch/dl_sniffer_pdsch.c")).read()
    m = re.search(r"dl_sniffer_tbs_format1c_table\[32\]\s*=\s*\{(.*?)\};", src, re.S)
    assert [int(x) for x in re.findall(r"\d+", m.group(1))] == f1c
    src = open(os.path.join(REF, "lib/src/phy/falcon_phch/ul_sniffer_pusch.c")).read()
    m = re.search(r"tbs_table_32A\[110\]\s*=\s*(?:/\*.*?\*/)?\s*\{(.*?)\};", src, re.S)
    row32a = np.array([int(x) for x in re.findall(r"\d+", m.group(1))])
    assert len(row32a) == 110
    # 32A lies between rows 32 and 33 of Table 7.1.7.2.1-1 and only uses transport block sizes of the table
    assert np.all(row32a >= T[31]) and np.all(row32a[:100] <= T[33][:100])   # (32A keeps growing to 101840 above 100 PRB)
    assert set(row32a.tolist()) <= set(T.ravel().tolist()) | {101840}
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lte_tables.h")).read()
    m = re.search(r"lte_tbs_row_32a\[LTE_TBS_NOF_PRB\]\s*=\s*\{(.*?)\};", hdr, re.S)
    assert [int(x) for x in re.findall(r"\d+", m.group(1))] == row32a.tolist()      # our copy of the row == the reference's


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_transport_block_sizes_of_the_references_captures_are_in_the_table():
    """Real-network anchor for the TBS table (36.213 Table 7.1.7.2.1-1, restated from the specification because srsRAN is absent): every MAC PDU in the
    reference's example captures (pcap_file_example/*.pcap, written from live cells by the reference) is a decoded transport block, so its size must be an
    entry of the table -- 86 distinct sizes, down- and uplink.  A random byte size hits the table with probability < 1/3 in that range."""
    import test_sinks
    _, _, tbs, f1c = check_tables.load()
    allT = set(np.array(tbs).ravel().tolist())
    sizes = set()
    for path in test_sinks.FILES:
        _, recs = test_sinks.parse(path)
        sizes |= {len(r["pdu"]) * 8 for r in recs}
    assert len(sizes) >= 80 and sizes <= allT, sorted(sizes - allT)
    lim = max(sizes)
    assert sum(1 for v in range(16, lim + 1, 8) if v in allT) * 3 < (lim - 8) // 8
