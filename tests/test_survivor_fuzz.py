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


@pytest.mark.parametrize("world,seed", [(2, 11), (3, 12), (8, 13)])
def test_packed_walk_of_a_sharded_batch_agrees_on_random_tables(infra, world, seed):
    """The walk turn of the sharded path (ltephy_pack_subframes per rank -> ltephy_search_batch_packed over all ranks' records, global subframe g = local
    index * world + rank) against the plain full-table walk in global order, on the adversarial tables above: same DCIs with global subframe indices, same
    (tti, cfi) table, same statistics; after a RAR activation ltephy_packed_needs_full_table says so on every rank's view and the walk refuses the records
    until the full tables come with them."""
    L = capi.load_library()
    capi._bind_search(L)
    rng = np.random.default_rng(seed)
    cellp = (50, 2, 21, 2)
    plain, packed = capi.Search(*cellp), capi.Search(*cellp)
    plain.config(1, 0, 7)
    packed.config(1, 0, 7)
    S = infra.sim()
    cell = Cell(*cellp)
    sizes_n = len({S.lte_dci_sizeof(C.byref(cell), f) for f in range(9)})
    o = ltelib.Oracle(cell)
    pool = rng.integers(0x100, 0xFFF0, 12)
    n_loc, tti0 = 6, 0
    for batch in range(6):
        if batch == 4:
            for s in (plain, packed):
                L.ltephy_search_activate(s.h, 0x7A7A, 0, 2)
            pool = np.append(pool, 0x7A7A)
        N = n_loc * world
        infos = (capi.SfInfo * N)()
        tables = np.zeros((N, capi.MAX_LOC, capi.MAX_SIZES), capi.CAND_DTYPE)
        for g in range(N):
            cfi = int(rng.integers(1, 4))
            nof_cce = int(infra.oracle().lteo_nof_cce(o.h, cfi))
            infos[g].tti, infos[g].cfi, infos[g].nof_cce = (tti0 + g) % 10240, cfi, nof_cce
            infos[g].snr_db = 20.0 if rng.random() > 0.05 else 3.0
            for pp in range(2):                                  # the packed header carries the sums the SNR is formed from
                for aa in range(2):
                    infos[g].noise[pp][aa], infos[g].rsrp[pp][aa] = 1.0, 100.0 if infos[g].snr_db > 6 else 2.0
            pw = np.where(rng.random(nof_cce) < 0.15, 0.3, 1.2).astype(np.float32)
            for c in range(nof_cce):
                infos[g].cce_power[c] = pw[c]
            tables[g] = _random_table(rng, L, nof_cce, infos[g].tti % 10, pool, sizes_n)
        tti0 += N
        want = []
        for g in range(N):
            for d in plain.subframe(infos[g], tables[g], max_out=256):
                want.append((g, int(d["rnti"]), int(d["format"]), int(d["L"]), int(d["ncce"]), int(d["nof_bits"]), int(d["bits"]), int(d["histogram_value"])))
        bufs, offs, fulls = [], [], []
        for r in range(world):
            li = (capi.SfInfo * n_loc)(*[infos[i * world + r] for i in range(n_loc)])
            comp = np.zeros(n_loc, capi.COMPACT_DTYPE)
            for i in range(n_loc):
                comp[i] = packed.compact_from_table(infos[i * world + r], tables[i * world + r])[0]
            rec, of = capi.pack_subframes(packed, li, comp)
            pad = np.zeros(n_loc * capi.PACK_MAX_BYTES, np.uint8)
            pad[:len(rec)] = rec
            bufs.append(pad), offs.append(np.ascontiguousarray(of))
            fulls.append(np.ascontiguousarray(tables[r::world]))
        bp = (C.c_void_p * world)(*[b.ctypes.data for b in bufs])
        op = (C.c_void_p * world)(*[x.ctypes.data for x in offs])
        need = L.ltephy_packed_needs_full_table(packed.h, bp, op, world, n_loc)
        over = any(int(packed.compact_from_table(infos[g], tables[g])["count"][0]) > capi.COMPACT_CAP for g in range(N))
        assert need == (1 if batch >= 4 or over else 0)
        pk = capi.search_batch_packed(packed, bufs, offs, n_loc, 256 * N)
        if need:
            assert pk is None                                    # refused as a whole, nothing consumed
            pk = capi.search_batch_packed(packed, bufs, offs, n_loc, 256 * N, full=fulls)
        dcis, tc = pk
        got = [(int(d["sf"]), int(d["rnti"]), int(d["format"]), int(d["L"]), int(d["ncce"]), int(d["nof_bits"]), int(d["bits"]), int(d["histogram_value"])) for d in dcis]
        assert got == want, (world, batch)
        assert [(int(a), int(b)) for a, b in tc] == [(infos[g].tti, infos[g].cfi) for g in range(N)]
    sa, sb = plain.stats(), packed.stats()
    assert (sa.nof_decoded_locations, sa.nof_cce, sa.nof_missed_cce, sa.nof_subframes, sa.nof_locations) == \
           (sb.nof_decoded_locations, sb.nof_cce, sb.nof_missed_cce, sb.nof_subframes, sb.nof_locations)
