"""BASELINE.json configs 3 and 4 at full size on the GPU, checked through size-independent properties
(pytest -m gpu): planted DCIs are recovered and identical inputs give identical tables over an 8192-subframe
sweep; 10 000 transport blocks' worth of K=5824 code blocks survive encode -> AWGN -> decode with CRC ok."""
import ctypes as C
import json
import os
import time
import numpy as np
import pytest
import ltelib
from ltelib import Cell
from ltesniffer_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _note(name, d):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", name), "w") as f:
        json.dump(d, f)


def test_config3_dci_sweep_8192_subframes(infra, phylib):
    S = infra.sim()
    cell = Cell(100, 2, 1, 1)
    NU, N = 256, 8192
    phy = capi.LtePhy(100, 2, 1, 1, max_subframes=N)
    sizes, sidx = phy.sizes()
    rng = np.random.default_rng(33)
    ncce = phy.nof_cce(3)
    llr_u = (0.3 * rng.standard_normal((NU, capi.LLR_STRIDE))).astype(np.float32)
    llr_u[:, 72 * ncce:] = 0
    planted = []
    for i in range(NU):
        c = 0
        while c < 84:
            L = int(rng.integers(0, 4))
            c = (c + (1 << L) - 1) // (1 << L) * (1 << L)
            if c + (1 << L) > 84:
                break
            if rng.random() < 0.3:
                f = int(rng.choice([0, 1, 2, 4, 6, 7]))   # formats 0/1A, 1, 1C, 2, 2A
                nb = sizes[f]
                if nb + 16 <= 0.8 * (72 << L):
                    b = rng.integers(0, 2, nb).astype(np.uint8)
                    if f == 0:
                        b[0] = 0
                    if f == 2:
                        b[0] = 1
                    rnti = int(rng.integers(1, 65535))
                    e = np.zeros(72 << L, np.uint8)
                    S.lte_sim_pdcch_encode(ltelib.ptr(b), nb, rnti, L, ltelib.ptr(e))
                    llr_u[i, 72 * c:72 * c + (72 << L)] = (2.0 * e - 1.0) + 0.3 * rng.standard_normal(72 << L)
                    planted.append((i, c, L, f, rnti, b))
            c += 1 << L
    llr = np.tile(llr_u, (N // NU, 1))
    t0 = time.time()
    cands = phy.dci_sweep(llr, np.full(N, 3, np.uint32))
    dt = time.time() - t0
    kern_ms = phy.timing()[3]
    nc, Ls = phy.locations(3)
    loc_of = {(int(nc[j]), int(Ls[j])): j for j in range(len(nc))}
    assert len(nc) == 157
    ndec = N * 157 * len(set(sidx))
    for (i, c, L, f, rnti, b) in planted:
        for rep in (0, N // NU - 1):
            cd = cands[i + rep * NU, loc_of[(c, L)], sidx[f]]
            assert cd["valid"] == 1 and int(cd["rnti"]) == rnti and np.array_equal(capi.cand_bits(cd["bits"], len(b)), b)
    # identical inputs -> identical tables (every tile equals the first)
    first = cands[:NU].view(np.uint8)
    for rep in range(1, N // NU):
        assert np.array_equal(cands[rep * NU:(rep + 1) * NU].view(np.uint8), first)
    assert len(planted) > 1000
    _note("config3_dci_sweep.json", {"subframes": N, "viterbi_decodes": ndec, "kernel_ms": kern_ms, "decodes_per_s": ndec / (kern_ms * 1e-3),
                                     "wall_s_with_copies": dt, "planted_recovered": len(planted)})
    phy.close()


def test_config4_turbo_10k_codewords(infra, phylib):
    S = infra.sim()
    S.lte_turbo_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    K, D = 5824, 5828            # TBS 75376 -> 13 code blocks of K = 5824
    NU, CHUNK_TB, NCHUNK = 1300, 1000, 10
    rng = np.random.default_rng(44)
    d_u = np.zeros((NU, 3 * D), np.int16)
    info_u = np.zeros((NU, K), np.uint8)
    for i in range(NU):
        b = rng.integers(0, 2, K).astype(np.uint8)
        crc = S.lte_crc(0x1800063, 24, ltelib.ptr(b), K - 24)
        for j in range(24):
            b[K - 24 + j] = (crc >> (23 - j)) & 1
        info_u[i] = b
        enc = np.zeros(3 * D, np.uint8)
        S.lte_turbo_encode(ltelib.ptr(b), K, ltelib.ptr(enc[0:]), ltelib.ptr(enc[D:]), ltelib.ptr(enc[2 * D:]))
        x = (2.0 * enc - 1.0) + 0.79 * rng.standard_normal(3 * D)    # Es/N0 = 2 dB at rate 1/3
        d_u[i] = np.clip(np.round(x * 48), -255, 255).astype(np.int16)
    phy = capi.LtePhy(100, 1, 1, 1, max_subframes=1)
    tot_ms = 0.0
    tot_bits = 0
    iters_hist = np.zeros(9, np.int64)
    for ch in range(NCHUNK):
        sel = rng.integers(0, NU, CHUNK_TB * 13)
        bits, iters, ok = phy.turbo_batch(d_u[sel], K, 8, 2)
        assert ok.all(), "chunk %d: %d code blocks failed CRC24B" % (ch, int((ok == 0).sum()))
        assert np.array_equal(bits, info_u[sel])
        iters_hist += np.bincount(iters, minlength=9)[:9]
        tot_ms += phy.timing()[2]
        tot_bits += CHUNK_TB * 75376
    # fixed 8 iterations (early stop off) for the throughput figure BASELINE.json names
    sel = rng.integers(0, NU, CHUNK_TB * 13)
    bits, iters, ok = phy.turbo_batch(d_u[sel], K, 8, 0)
    fixed_ms = phy.timing()[2]
    assert np.array_equal(bits, info_u[sel]) and (iters == 8).all()
    _note("config4_turbo.json", {"codewords": NCHUNK * CHUNK_TB, "code_blocks": NCHUNK * CHUNK_TB * 13, "K": K,
                                 "early_stop_mbit_s": tot_bits / tot_ms / 1e3, "iters_hist": iters_hist.tolist(),
                                 "fixed8_mbit_s": CHUNK_TB * 75376 / fixed_ms / 1e3, "fixed8_kernel_ms_per_1000_codewords": fixed_ms})
    phy.close()
