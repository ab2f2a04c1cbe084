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


def gather_tables(info_local, cands_local, world, device):
    """info_local: ctypes (SfInfo * B); cands_local: uint8 tensor [B, MAX_LOC, MAX_SIZES, 16] (host).
    Returns (info_all ctypes array [B*world] in global order, cands_all uint8 host tensor [B*world, ...])."""
    B = len(info_local)
    isz = C.sizeof(capi.SfInfo)
    li = torch.frombuffer(info_local, dtype=torch.uint8).clone().to(device)
    lc = cands_local.reshape(-1).to(device)
    gi = [torch.empty_like(li) for _ in range(world)]
    gc = [torch.empty_like(lc) for _ in range(world)]
    dist.all_gather(gi, li)
    dist.all_gather(gc, lc)
    ia = torch.stack(gi).view(world, B, isz).transpose(0, 1).contiguous().cpu().numpy()   # [B][world] -> g = i*world + r
    info_all = (capi.SfInfo * (B * world))()
    C.memmove(info_all, ia.ctypes.data, ia.nbytes)
    ca = torch.stack(gc).view(world, B, capi.MAX_LOC, capi.MAX_SIZES, 16).transpose(0, 1).reshape(B * world, capi.MAX_LOC, capi.MAX_SIZES, 16).contiguous().cpu()
    return info_all, ca


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
