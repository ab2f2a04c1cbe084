"""Round-robin subframe sharding across ranks: global subframe g is owned by rank g % world.
Every rank runs phase A on its own subframes; the per-subframe records (ltephy_sf_info_t) and candidate tables
are all-gathered, re-interleaved into global order, and the FALCON walk is replayed over ALL subframes on every
rank (its RNTI history is inherently sequential); each rank then keeps the grants of the subframes it owns.
Host logic only: works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist
from . import capi


_stage = {}


def _buffers(B, world, device):
    key = (B, world, str(device))
    if key not in _stage:
        isz = C.sizeof(capi.SfInfo)
        csz = capi.MAX_LOC * capi.MAX_SIZES * 16
        pin = torch.device(device).type == "cuda"
        _stage[key] = dict(
            gi=torch.empty((world, B, isz), dtype=torch.uint8, device=device), gc=torch.empty((world, B, csz), dtype=torch.uint8, device=device),
            oi=torch.empty((B, world, isz), dtype=torch.uint8, device=device), oc=torch.empty((B, world, csz), dtype=torch.uint8, device=device),
            hi=torch.empty((B * world, isz), dtype=torch.uint8, pin_memory=pin), hc=torch.empty((B * world, capi.MAX_LOC, capi.MAX_SIZES, 16), dtype=torch.uint8, pin_memory=pin),
            info_all=(capi.SfInfo * (B * world))())
    return _stage[key]


def gather_tables(info_local, cands_local, world, device):
    """info_local: ctypes (SfInfo * B); cands_local: uint8 tensor [B, MAX_LOC, MAX_SIZES, 16] (host, pinned if possible).
    Returns (info_all ctypes array [B*world] in global order g = i*world + r, cands_all uint8 host tensor [B*world, ...]).
    The returned buffers are reused by the next call with the same shape."""
    B = len(info_local)
    st = _buffers(B, world, device)
    li = torch.frombuffer(info_local, dtype=torch.uint8).to(device, non_blocking=True)
    lc = cands_local.reshape(-1).to(device, non_blocking=True)
    dist.all_gather_into_tensor(st["gi"].view(-1), li)
    dist.all_gather_into_tensor(st["gc"].view(-1), lc)
    st["oi"].copy_(st["gi"].transpose(0, 1))          # [world][B] -> [B][world]
    st["oc"].copy_(st["gc"].transpose(0, 1))
    st["hi"].copy_(st["oi"].view(B * world, -1), non_blocking=True)
    st["hc"].view(B * world, -1).copy_(st["oc"].view(B * world, -1), non_blocking=True)
    if torch.device(device).type == "cuda":
        torch.cuda.current_stream().synchronize()
    C.memmove(st["info_all"], st["hi"].data_ptr(), st["hi"].numel())
    return st["info_all"], st["hc"]


def search_and_select(L, srch, info_all, cands_all, world, rank, max_dcis, max_grants):
    """walk over all subframes in global order; -> (dcis structured array, grants ctypes array, grant->dci index, n_grants)"""
    capi._bind_search(L)
    n = len(info_all)
    dcis = np.zeros(max_dcis, capi.DCI_DTYPE)
    nd = C.c_uint32(0)
    r = L.ltephy_search_batch(srch.h, info_all, C.c_void_p(cands_all.data_ptr()), n, dcis.ctypes.data_as(C.c_void_p), max_dcis, C.byref(nd))
    if r != 0:
        raise RuntimeError("ltephy_search_batch failed (%d)" % r)
    grants = (capi.Grant * max_grants)()
    gidx = np.zeros(max_grants, np.uint32)
    ng = C.c_uint32(0)
    r = L.ltephy_grants_from_dcis(srch.h, info_all, dcis.ctypes.data_as(C.c_void_p), nd.value, world, rank, grants, gidx.ctypes.data_as(C.c_void_p),
                                  max_grants, C.byref(ng))
    if r != 0:
        raise RuntimeError("ltephy_grants_from_dcis failed (%d)" % r)
    return dcis[:nd.value], grants, gidx, ng.value
