"""Randomised equivalence of the two forms of the FALCON walk: ltephy_search_subframe over the full candidate table and
ltephy_search_subframe_compact over its survivor form (ltephy_compact_from_table) must report the same DCIs, histogram values
and statistics on tables built to hit the corner cases: RNTIs drawn from a small pool (so histograms cross their threshold and
RNTIs become active), entries repeated in first children (shortcut), zero RNTIs, undecoded entries, low-power CCEs, all three
CFIs, meta-format re-splits, and RAR-activated RNTIs (where the survivor form must be refused)."""
import ctypes as C
import numpy as np
import pytest
import ltelib
from ltelib import Cell
from ltesniffer_b200 import capi


def _locations(nof_cce):
    lim = min(nof_cce, 84)
    out = []
    for l in (3, 2, 1, 0):
        for i in range(lim >> l):
            out.append((l, i << l))
    return out


def _random_table(rng, L, nof_cce, sf_idx, pool, sizes_n):
    """-> CAND_DTYPE [MAX_LOC][MAX_SIZES] with structure"""
    T = np.zeros((capi.MAX_LOC, capi.MAX_SIZES), capi.CAND_DTYPE)
    locs = _locations(nof_cce)
    index = {lc: i for i, lc in enumerate(locs)}
    for li, (l, ncce) in enumerate(locs):
        for si in range(sizes_n):
            u = rng.random()
            if u < 0.03:
                continue                                         # undecoded entry (all-zero LLRs)
            if u < 0.06:
                r = 0                                            # decoded RNTI 0
            elif u < 0.45:
                r = int(rng.choice(pool))                        # a plausible RNTI (may or may not fit the location)
            elif u < 0.50:
                r = int(rng.choice([1, 5, 10, 0xFFFE, 0xFFFF, 0xFFFD, 0xFFF4]))   # RA / paging / SI / reserved ranges
            else:
                r = int(rng.integers(1, 65536))                  # noise
            T[li, si] = (int(rng.integers(0, 1 << 63)) << 1 | int(rng.integers(0, 2)), r, 1, [0] * 5)
    # plant "real" DCIs: same RNTI and bits at a location of the RNTI's search space and in its first children
    for _ in range(int(rng.integers(2, 9))):
        r = int(rng.choice(pool))
        l = int(rng.integers(0, 4))
        cands = [nc for (ll, nc) in locs if ll == l and L.ltephy_search_validate_location(nof_cce, nc, l, sf_idx, r)]
        if not cands:
            continue
        ncce = int(rng.choice(cands))
        si = int(rng.integers(0, sizes_n))
        bits = int(rng.integers(0, 1 << 63)) << 1 | int(rng.integers(0, 2))
        ll, nc = l, ncce
        while True:
            T[index[(ll, nc)], si] = (bits, r, 1, [0] * 5)
            if ll == 0 or rng.random() < 0.3:
                break
            ll -= 1                                              # first child keeps the CCE index
    return T


@pytest.mark.parametrize("cellp,seed", [((100, 2, 7, 2), 1), ((50, 1, 3, 1), 2), ((25, 2, 11, 2), 3), ((75, 2, 200, 1), 4)])
def test_full_and_survivor_walks_agree_on_random_tables(infra, cellp, seed):
    L = capi.load_library()
    capi._bind_search(L)
    rng = np.random.default_rng(seed)
    full, comp = capi.Search(*cellp), capi.Search(*cellp)
    full.config(1, 0, 7)
    comp.config(1, 0, 7)                                         # re-split the meta formats every 7 subframes
    S = infra.sim()
    cell = Cell(*cellp)
    sizes_n = len({S.lte_dci_sizeof(C.byref(cell), f) for f in range(9)})    # size columns of the candidate table
    o = ltelib.Oracle(cell)
    pool = rng.integers(0x100, 0xFFF0, 12)
    ndci = 0
    refused = 0
    for sf in range(260):
        cfi = int(rng.integers(1, 4))
        nof_cce = int(infra.oracle().lteo_nof_cce(o.h, cfi))
        info = capi.SfInfo()
        info.tti, info.cfi, info.nof_cce = sf, cfi, nof_cce
        info.snr_db = 20.0 if rng.random() > 0.03 else 3.0       # a few subframes fail the 6 dB gate
        pw = np.where(rng.random(nof_cce) < 0.15, 0.3, 1.2).astype(np.float32)
        for c in range(nof_cce):
            info.cce_power[c] = pw[c]
        if sf == 200:                                            # from here on: a fresh RAR-activated RNTI on both histories,
            rar = 0x7A7A                                         # which then also shows up in the tables
            for s in (full, comp):
                L.ltephy_search_activate(s.h, rar, 0, 2)
            pool = np.append(pool, rar)
        T = _random_table(rng, L, nof_cce, sf % 10, pool, sizes_n)
        a = full.subframe(info, T, max_out=256)
        cf = comp.compact_from_table(info, T)
        b = comp.subframe_compact(info, cf, max_out=256)
        if b is None:                                            # refused: nothing consumed, the full table must be used
            refused += 1
            assert sf >= 200 or int(cf["count"][0]) > capi.COMPACT_CAP
            b = comp.subframe(info, T, max_out=256)
        assert len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in a.dtype.names), (cellp, sf)
        ndci += len(a)
    sa, sb = full.stats(), comp.stats()
    assert (sa.nof_decoded_locations, sa.nof_cce, sa.nof_missed_cce, sa.nof_subframes, sa.nof_locations) == \
           (sb.nof_decoded_locations, sb.nof_cce, sb.nof_missed_cce, sb.nof_subframes, sb.nof_locations)
    assert ndci > 200 and refused >= 1
