"""Multi-rank sharding logic on CPU (gloo, world_size 2): candidate tables built per rank for the subframes it
owns (with the CPU oracle standing in for the GPU phase A, and the host restatement of the survivor selection
standing in for cand_compact_kernel), all-gathered and re-interleaved; the walk replayed on every rank must accept
exactly what a single process walking the FULL tables accepts, and the grants must partition by owner.  A second
pass with a RAR-activated RNTI forces the full-table fallback on both ranks."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_SF = 8


def _tables(cell_args, ttis):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import ltelib
    from ltesniffer_b200 import capi
    from test_host_search import oracle_table, host_geometry, locations
    cell = ltelib.Cell(*cell_args)
    s = ltelib.Sim(cell=cell, seed=9, cfi=2, nof_ues=4, dl_min=2, dl_max=3, tm=1, mcs_min=3, mcs_max=9, snr_db=27.0, fixed_L=2, si_period=4)
    o = ltelib.Oracle(cell)
    geo = host_geometry(cell)
    info = (capi.SfInfo * len(ttis))()
    cands = np.zeros((len(ttis), capi.MAX_LOC, capi.MAX_SIZES), capi.CAND_DTYPE)
    for i, tti in enumerate(ttis):
        iq, tr, pl = s.subframe(int(tti))
        sym = o.ofdm(iq)
        ce, res = o.chest(int(tti) % 10, sym)
        cfi, _ = o.pcfich(int(tti) % 10, sym, ce)
        llr = o.pdcch_llr(int(tti) % 10, cfi, sym, ce)
        ncce = len(llr) // 72
        info[i].tti, info[i].cfi, info[i].nof_cce, info[i].snr_db = int(tti), cfi, ncce, res.snr_db
        for pp in range(2):                      # per-path sums: the packed exchange format derives snr_db from them
            for aa in range(2):
                info[i].noise[pp][aa], info[i].rsrp[pp][aa] = res.noise[pp][aa], res.rsrp[pp][aa]
        pw = np.zeros(ncce, np.float32)
        ltelib.oracle().lteo_cce_power(ltelib.ptr(llr), ncce, ltelib.ptr(pw))
        for c in range(ncce):
            info[i].cce_power[c] = pw[c]
        nc, Ls = locations(ncce)
        cands[i] = oracle_table(o, geo, nc, Ls, llr)
    return cell, info, cands


def _compact(srch, info, cands):
    from ltesniffer_b200 import capi
    comp = np.zeros(len(info), capi.COMPACT_DTYPE)
    for i in range(len(info)):
        comp[i] = srch.compact_from_table(info[i], cands[i])[0]
    return torch.from_numpy(comp.view(np.uint8).reshape(len(info), capi.COMPACT_DTYPE.itemsize))


RAR_RNTI = 0x4321


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import ctypes as C
    from ltesniffer_b200 import capi, shard
    cell_args = (50, 1, 3, 1)
    mine = np.arange(rank, N_SF, world)
    cell, info, cands = _tables(cell_args, mine)
    ct = torch.from_numpy(cands.view(np.uint8).reshape(len(mine), capi.MAX_LOC, capi.MAX_SIZES, 16))
    L = capi.load_library()
    srch = capi.Search(*cell_args)
    info_all, comp_all = shard.gather_tables(info, _compact(srch, info, cands), world, "cpu")
    assert [info_all[g].tti for g in range(N_SF)] == list(range(N_SF))
    dcis, grants, gidx, ng = shard.search_and_select(L, srch, info_all, comp_all, world, rank, 64 * N_SF, 64 * N_SF)
    # second pass on a fresh history with a RAR-activated RNTI: the survivor form must be refused and the full tables gathered
    srch2 = capi.Search(*cell_args)
    L.ltephy_search_activate(srch2.h, RAR_RNTI, 0, 2)
    assert not shard.need_full_tables(L, srch, comp_all, N_SF) and shard.need_full_tables(L, srch2, comp_all, N_SF)
    fetched = []

    def full_fetch():
        fetched.append(1)
        return shard.gather_full_tables(ct, world, "cpu")
    dcis2, _, _, _ = shard.search_and_select(L, srch2, info_all, comp_all, world, rank, 64 * N_SF, 64 * N_SF, full_fetch)
    key = lambda d: (int(d["sf"]), int(d["rnti"]), int(d["format"]), int(d["ncce"]), int(d["L"]), int(d["bits"]))
    # ---- the packed exchange format of ltephy_decode_subframes_sharded (include/ltephy_shard.h): records packed to their used
    # length per rank, offsets and records all-gathered, the walk reads them in place
    comp_np = _compact(srch, info, cands).numpy().view(capi.COMPACT_DTYPE).reshape(-1)
    rec, offs = capi.pack_subframes(srch, info, comp_np)
    n_loc = len(mine)
    assert offs[n_loc] == len(rec) < n_loc * capi.PACK_MAX_BYTES and all(int(o) % 16 == 0 for o in offs)
    offs_all = [torch.zeros(n_loc + 1, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(offs_all, torch.from_numpy(offs.view(np.int32)))
    pad = torch.zeros(n_loc * capi.PACK_MAX_BYTES, dtype=torch.uint8)
    pad[:len(rec)] = torch.from_numpy(rec)
    rec_all = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(rec_all, pad)
    bufs = [np.ascontiguousarray(t.numpy()) for t in rec_all]
    offl = [np.ascontiguousarray(t.numpy().view(np.uint32)) for t in offs_all]
    srch3 = capi.Search(*cell_args)
    pk = capi.search_batch_packed(srch3, bufs, offl, n_loc, 64 * N_SF)
    assert pk is not None
    dcis3, tc = pk
    assert [int(x) for x in tc[:, 0]] == list(range(N_SF))
    gr3 = (capi.Grant * (64 * N_SF))()
    gi3 = np.zeros(64 * N_SF, np.uint32)
    ng3 = C.c_uint32(0)
    assert L.ltephy_grants_from_dcis_tc(srch3.h, tc.ctypes.data_as(C.c_void_p), dcis3.ctypes.data_as(C.c_void_p), len(dcis3), world, rank, gr3,
                                        gi3.ctypes.data_as(C.c_void_p), 64 * N_SF, C.byref(ng3)) == 0
    srch4 = capi.Search(*cell_args)
    L.ltephy_search_activate(srch4.h, RAR_RNTI, 0, 2)
    assert capi.search_batch_packed(srch4, bufs, offl, n_loc, 64 * N_SF) is None      # refused, nothing consumed
    full_all = shard.gather_full_tables(ct, world, "cpu").numpy().reshape(n_loc, world, -1)   # [i][r] global order -> per rank
    fulls = [np.ascontiguousarray(full_all[:, r]) for r in range(world)]
    dcis4, _ = capi.search_batch_packed(srch4, bufs, offl, n_loc, 64 * N_SF, full=fulls)
    q.put((rank, [key(d) for d in dcis], [(int(grants[i].sf), int(grants[i].rnti), int(grants[i].nof_re), int(gidx[i])) for i in range(ng)],
           [key(d) for d in dcis2], len(fetched), [key(d) for d in dcis3],
           [(int(gr3[i].sf), int(gr3[i].rnti), int(gr3[i].nof_re), int(gi3[i])) for i in range(ng3.value)], [key(d) for d in dcis4]))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(infra):
    sys.path.insert(0, ROOT)
    import ctypes as C
    from ltesniffer_b200 import capi, shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=600) for _ in range(2)])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process reference: one history, FULL tables
    cell_args = (50, 1, 3, 1)
    cell, info, cands = _tables(cell_args, np.arange(N_SF))
    L = capi.load_library()
    capi._bind_search(L)
    key = lambda d: (int(d["sf"]), int(d["rnti"]), int(d["format"]), int(d["ncce"]), int(d["L"]), int(d["bits"]))

    def walk_full(srch):
        dcis = np.zeros(64 * N_SF, capi.DCI_DTYPE)
        nd = C.c_uint32(0)
        assert L.ltephy_search_batch(srch.h, info, cands.ctypes.data_as(C.c_void_p), N_SF, dcis.ctypes.data_as(C.c_void_p), len(dcis), C.byref(nd)) == 0
        return [key(d) for d in dcis[:nd.value]]
    ref_dcis = walk_full(capi.Search(*cell_args))
    assert len(ref_dcis) >= N_SF // 2
    assert out[0][1] == ref_dcis and out[1][1] == ref_dcis           # every rank accepts the same DCIs as one process
    srch2 = capi.Search(*cell_args)
    L.ltephy_search_activate(srch2.h, RAR_RNTI, 0, 2)
    ref_dcis2 = walk_full(srch2)
    assert all(o[3] == ref_dcis2 and o[4] == 1 for o in out)         # fallback taken once per rank, same result as one process
    # packed exchange format: same DCIs, same grants, same fallback result
    assert all(o[5] == ref_dcis and o[6] == o[2] and o[7] == ref_dcis2 for o in out)
    srch3 = capi.Search(*cell_args)
    dcis, grants, gidx, ng = shard.search_and_select(L, srch3, info, _compact(srch3, info, cands), 1, 0, 64 * N_SF, 64 * N_SF)
    assert [key(d) for d in dcis] == ref_dcis
    ref_grants = [(int(grants[i].sf), int(grants[i].rnti), int(grants[i].nof_re), int(gidx[i])) for i in range(ng)]
    merged = sorted([(g[0] * 2 + r, g[1], g[2], g[3]) for r in range(2) for g in out[r][2]], key=lambda x: x[3])
    assert merged == sorted(ref_grants, key=lambda x: x[3])          # grants partition by owner, local sf = g // world
