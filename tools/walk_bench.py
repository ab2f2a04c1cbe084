"""Host-only timing of the FALCON walk (ltephy_search_batch vs ltephy_search_batch_compact) on the bench workload.
Tables come from the CPU oracle (tests/ltelib.py), so this runs without a GPU.  usage: python tools/walk_bench.py [n_unique]"""
import ctypes as C
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
import ltelib
from ltesniffer_b200 import capi
from test_host_search import oracle_table, host_geometry, locations

nu = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cell = ltelib.Cell(bench.CELL["nof_prb"], bench.CELL["nof_ports"], bench.CELL["cell_id"], bench.CELL["nof_rx"])
s, o = ltelib.Sim(cell=cell, **bench.SIM_KW), ltelib.Oracle(cell)
geo = host_geometry(cell)
srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
L = srch.L
info = (capi.SfInfo * nu)()
cands = np.zeros((nu, capi.MAX_LOC, capi.MAX_SIZES), capi.CAND_DTYPE)
comp = np.zeros(nu, capi.COMPACT_DTYPE)
for tti in range(nu):
    iq, tr, pl = s.subframe(tti)
    sym = o.ofdm(iq)
    ce, res = o.chest(tti % 10, sym)
    cfi, corr = o.pcfich(tti % 10, sym, ce)
    llr = o.pdcch_llr(tti % 10, cfi, sym, ce)
    ncce = len(llr) // 72
    info[tti].tti, info[tti].cfi, info[tti].nof_cce, info[tti].snr_db = tti, cfi, ncce, res.snr_db
    pw = np.zeros(ncce, np.float32)
    ltelib.oracle().lteo_cce_power(ltelib.ptr(llr), ncce, ltelib.ptr(pw))
    for i in range(ncce):
        info[tti].cce_power[i] = pw[i]
    nc, Ls = locations(ncce)
    cands[tti] = oracle_table(o, geo, nc, Ls, llr)
    comp[tti] = srch.compact_from_table(info[tti], cands[tti])[0]
N = 1000
idx = np.arange(N) % nu
info_b = (capi.SfInfo * N)(*[info[i] for i in idx])
cands_b = np.ascontiguousarray(cands[idx])
comp_b = np.ascontiguousarray(comp[idx])
dcis = np.zeros(24 * N, capi.DCI_DTYPE)
nd = C.c_uint32(0)
for name in ("full", "compact"):
    sr = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
    ts = []
    for rep in range(8):
        t0 = time.perf_counter()
        if name == "full":
            r = L.ltephy_search_batch(sr.h, info_b, cands_b.ctypes.data_as(C.c_void_p), N, dcis.ctypes.data_as(C.c_void_p), len(dcis), C.byref(nd))
        else:
            r = L.ltephy_search_batch_compact(sr.h, info_b, comp_b.ctypes.data_as(C.c_void_p), None, N, dcis.ctypes.data_as(C.c_void_p), len(dcis), C.byref(nd))
        ts.append(time.perf_counter() - t0)
        assert r == 0
    print("%-8s %d subframes: best %.3f ms (%.2f us/sf), median %.3f ms, %d DCIs, survivors/sf %.1f" %
          (name, N, min(ts) * 1e3, min(ts) * 1e6 / N, sorted(ts)[len(ts) // 2] * 1e3, nd.value, comp["count"].mean()))
