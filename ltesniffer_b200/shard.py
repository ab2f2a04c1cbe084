"""Round-robin subframe sharding across ranks: global subframe g is owned by rank g % world.
Every rank runs phase A on its own subframes; the per-subframe records (ltephy_sf_info_t) and the survivor forms of the
candidate tables (ltephy_compact_t, 4.5 KB per subframe) are all-gathered, re-interleaved into global order, and the
FALCON walk is replayed over ALL subframes on every rank (its RNTI history is inherently sequential); each rank then
keeps the grants of the subframes it owns.  When the survivor form cannot serve the walk (RAR-activated RNTIs, an
overfull subframe) every rank sees that on the same data and the full tables are all-gathered instead.
Host logic only: works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist
from . import capi


_stage = {}
ISZ = C.sizeof(capi.SfInfo)
CSZ = capi.COMPACT_DTYPE.itemsize
FSZ = capi.MAX_LOC * capi.MAX_SIZES * 16


def _buffers(B, world, device, esz, slot=0):
    key = (B, world, str(device), esz, slot)
    if key not in _stage:
        pin = torch.device(device).type == "cuda"
        _stage[key] = dict(g=torch.empty((world, B, esz), dtype=torch.uint8, device=device), o=torch.empty((B, world, esz), dtype=torch.uint8, device=device),
                           h=torch.empty((B * world, esz), dtype=torch.uint8, pin_memory=pin))
    return _stage[key]


def _gather_interleaved(local_u8, B, world, device, esz, slot=0):
    """local_u8: uint8 tensor [B*esz] (host, or already on `device`) -> host tensor [B*world, esz] in global order
    g = i*world + r.  The result lives in a staging buffer owned by `slot` and is overwritten by its next call."""
    st = _buffers(B, world, device, esz, slot)
    loc = local_u8.reshape(-1)
    if loc.is_cuda != (torch.device(device).type == "cuda"):
        loc = loc.to(device, non_blocking=True)
    dist.all_gather_into_tensor(st["g"].view(-1), loc)
    st["o"].copy_(st["g"].transpose(0, 1))            # [world][B] -> [B][world]
    st["h"].copy_(st["o"].view(B * world, esz), non_blocking=True)
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream().synchronize()
    return st["h"]


_info_all = {}


def gather_tables(info_local, comp_local, world, device, slot=0):
    """info_local: ctypes (SfInfo * B); comp_local: uint8 tensor [B, sizeof(ltephy_compact_t)] (host, pinned if possible).
    -> (info_all ctypes array [B*world] in global order, comp_all uint8 host tensor [B*world, sizeof(ltephy_compact_t)]).
    The returned buffers belong to `slot` (one per pipeline) and are reused by its next call with the same shape."""
    B = len(info_local)
    hi = _gather_interleaved(torch.frombuffer(info_local, dtype=torch.uint8), B, world, device, ISZ, slot)
    hc = _gather_interleaved(comp_local, B, world, device, CSZ, slot)
    if (B, world, slot) not in _info_all:
        _info_all[(B, world, slot)] = (capi.SfInfo * (B * world))()
    info_all = _info_all[(B, world, slot)]
    C.memmove(info_all, hi.data_ptr(), hi.numel())
    return info_all, hc


_dev_local = {}


def gather_tables_device(L, phy_handle, B, world, nof_ports, nof_rx, slot=0):
    """CUDA path: the handle's raw per-subframe records and survivor forms go device -> device -> NCCL all-gather -> host,
    with no host->device copy (which would queue behind the IQ transfers of the other pipelines).  Same result as
    gather_tables(); the records are finalised (snr_db, cfo) on the host after the exchange."""
    key = (B, slot)
    if key not in _dev_local:
        _dev_local[key] = (torch.empty(B * ISZ, dtype=torch.uint8, device="cuda"), torch.empty(B * CSZ, dtype=torch.uint8, device="cuda"))
    di, dc = _dev_local[key]
    r = L.ltephy_copy_phase_a_device(phy_handle, C.c_void_p(di.data_ptr()), C.c_void_p(dc.data_ptr()))
    if r != 0:
        raise RuntimeError("ltephy_copy_phase_a_device failed (%d)" % r)
    hi = _gather_interleaved(di, B, world, "cuda", ISZ, slot)
    hc = _gather_interleaved(dc, B, world, "cuda", CSZ, slot)
    if (B, world, slot) not in _info_all:
        _info_all[(B, world, slot)] = (capi.SfInfo * (B * world))()
    info_all = _info_all[(B, world, slot)]
    C.memmove(info_all, hi.data_ptr(), hi.numel())
    L.ltephy_finalize_info(info_all, B * world, nof_ports, nof_rx)
    return info_all, hc


def gather_full_tables(cands_local, world, device, slot=0):
    """cands_local: uint8 tensor [B, MAX_LOC, MAX_SIZES, 16] -> host tensor [B*world, MAX_LOC*MAX_SIZES*16] in global order"""
    return _gather_interleaved(cands_local, cands_local.shape[0], world, device, FSZ, slot)


class WalkBuffers:
    """output buffers of search_and_select, allocated once per pipeline"""

    def __init__(self, max_dcis, max_grants):
        self.dcis = np.zeros(max_dcis, capi.DCI_DTYPE)
        self.grants = (capi.Grant * max_grants)()
        self.gidx = np.zeros(max_grants, np.uint32)


def need_full_tables(L, srch, comp_all, n):
    """True iff the walk over these survivor forms would be refused (decided before anything is consumed; every rank
    sees the same data and history, so every rank decides the same)."""
    return L.ltephy_search_needs_full_table(srch.h, C.c_void_p(comp_all.data_ptr()), n) != 0


def search_and_select(L, srch, info_all, comp_all, world, rank, max_dcis, max_grants, full_fetch=None, full=None, bufs=None):
    """walk over all subframes in global order; -> (dcis structured array, grants ctypes array, grant->dci index, n_grants).
    full: all-gathered FULL tables when already fetched; else full_fetch: callable returning them on demand (collective:
    every rank calls it on the same condition)."""
    capi._bind_search(L)
    n = len(info_all)
    bufs = bufs or WalkBuffers(max_dcis, max_grants)
    dcis, grants, gidx = bufs.dcis, bufs.grants, bufs.gidx
    nd = C.c_uint32(0)
    if full is not None:
        r = L.ltephy_search_batch_compact(srch.h, info_all, C.c_void_p(comp_all.data_ptr()), C.c_void_p(full.data_ptr()), n,
                                          dcis.ctypes.data_as(C.c_void_p), max_dcis, C.byref(nd))
        if r != 0:
            raise RuntimeError("ltephy_search_batch_compact failed (%d)" % r)
        return _select(L, srch, info_all, dcis, nd, world, rank, grants, gidx, max_grants)
    r = L.ltephy_search_batch_compact(srch.h, info_all, C.c_void_p(comp_all.data_ptr()), None, n, dcis.ctypes.data_as(C.c_void_p), max_dcis, C.byref(nd))
    if r == capi.NEED_FULL_TABLE:
        if full_fetch is None:
            raise RuntimeError("the survivor form cannot serve this batch and no full-table fetch was given")
        full = full_fetch()
        r = L.ltephy_search_batch_compact(srch.h, info_all, C.c_void_p(comp_all.data_ptr()), C.c_void_p(full.data_ptr()), n,
                                          dcis.ctypes.data_as(C.c_void_p), max_dcis, C.byref(nd))
    if r != 0:
        raise RuntimeError("ltephy_search_batch_compact failed (%d)" % r)
    return _select(L, srch, info_all, dcis, nd, world, rank, grants, gidx, max_grants)


def _select(L, srch, info_all, dcis, nd, world, rank, grants, gidx, max_grants):
    ng = C.c_uint32(0)
    r = L.ltephy_grants_from_dcis(srch.h, info_all, dcis.ctypes.data_as(C.c_void_p), nd.value, world, rank, grants, gidx.ctypes.data_as(C.c_void_p),
                                  max_grants, C.byref(ng))
    if r != 0:
        raise RuntimeError("ltephy_grants_from_dcis failed (%d)" % r)
    return dcis[:nd.value], grants, gidx, ng.value
