"""ctypes binding of include/ltephy_b200.h (the tier-1 C-ABI).  Host arrays in, host arrays out;
all compute happens in the CUDA library.  Fails loudly if the library is missing or unusable."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libltephy_b200.so")

MAX_PRB, MAX_CCE, MAX_LOC, MAX_SIZES, NOF_FORMATS = 110, 88, 160, 8, 9
LLR_STRIDE = 72 * MAX_CCE
TX_PORT0, TX_DIVERSITY, TX_CDD, TX_SPATIALMUX = 0, 1, 2, 3
TAP_SYM, TAP_CE, TAP_LLR, TAP_PDSCH_LLR, TAP_TURBO_IN = 0, 1, 2, 3, 4
FLAG_SKIP_LOW_POWER = 1


class Cfg(C.Structure):
    _fields_ = [("nof_prb", C.c_uint32), ("nof_ports", C.c_uint32), ("cell_id", C.c_uint32), ("nof_rx", C.c_uint32),
                ("max_subframes", C.c_uint32), ("max_grants", C.c_uint32), ("turbo_max_iter", C.c_uint32), ("device", C.c_int32),
                ("flags", C.c_uint32), ("symbol_sz", C.c_uint32), ("phich_resources", C.c_uint32), ("phich_length", C.c_uint32), ("reserved", C.c_uint32 * 4)]


class SfInfo(C.Structure):
    _fields_ = [("tti", C.c_uint32), ("cfi", C.c_uint32), ("nof_cce", C.c_uint32), ("nof_locations", C.c_uint32),
                ("pcfich_corr", C.c_float * 3), ("noise", (C.c_float * 2) * 2), ("rsrp", (C.c_float * 2) * 2),
                ("noise_avg", C.c_float), ("rsrp_avg", C.c_float), ("cfo_re", C.c_float), ("cfo_im", C.c_float),
                ("snr_db", C.c_float), ("cfo", C.c_float), ("rb_power", C.c_float * MAX_PRB), ("cce_power", C.c_float * MAX_CCE)]


class Cand(C.Structure):
    _fields_ = [("bits", C.c_uint64), ("rnti", C.c_uint16), ("valid", C.c_uint8), ("pad", C.c_uint8 * 5)]


class GrantTb(C.Structure):
    _fields_ = [("tbs", C.c_int32), ("qm", C.c_uint8), ("rv", C.c_uint8), ("enabled", C.c_uint8), ("cw_idx", C.c_uint8),
                ("harq_op", C.c_uint8), ("pad", C.c_uint8 * 3), ("harq_slot", C.c_uint32)]


class Grant(C.Structure):
    _fields_ = [("sf", C.c_uint32), ("rnti", C.c_uint16), ("tx_scheme", C.c_uint8), ("nof_tb", C.c_uint8),
                ("prb_mask", (C.c_uint32 * 4) * 2), ("nof_re", C.c_uint32), ("pmi", C.c_uint32), ("tb", GrantTb * 2)]


class TbResult(C.Structure):
    _fields_ = [("crc", C.c_uint8), ("avg_iters", C.c_uint8), ("nof_cb", C.c_uint16), ("payload_off", C.c_uint32), ("payload_len", C.c_uint32)]


class UlCfg(C.Structure):
    _fields_ = [("n_dmrs1", C.c_uint32), ("delta_ss", C.c_uint32), ("group_hopping", C.c_uint32), ("seq_hopping", C.c_uint32)]


class UlGrant(C.Structure):
    _fields_ = [("sf", C.c_uint32), ("rnti", C.c_uint16), ("qm", C.c_uint8), ("rv", C.c_uint8), ("L_prb", C.c_uint32), ("n_prb", C.c_uint32),
                ("n_dmrs2", C.c_uint32), ("tbs", C.c_int32), ("n_prb_slot1", C.c_uint32), ("nof_ack", C.c_uint8), ("ri_len", C.c_uint8),
                ("cqi_len", C.c_uint16), ("I_offset_ack", C.c_uint8), ("I_offset_cqi", C.c_uint8), ("I_offset_ri", C.c_uint8), ("flags", C.c_uint8)]


class UlUeCfg(C.Structure):
    _fields_ = [("rnti", C.c_uint16), ("mcs_mod", C.c_uint8), ("I_offset_ack", C.c_uint8), ("I_offset_cqi", C.c_uint8), ("I_offset_ri", C.c_uint8),
                ("cqi_len", C.c_uint16)]


class Rar(C.Structure):
    _fields_ = [("t_crnti", C.c_uint16), ("ta", C.c_uint16), ("rapid", C.c_uint8), ("hopping_flag", C.c_uint8), ("tpc", C.c_uint8), ("ul_delay", C.c_uint8),
                ("cqi_request", C.c_uint8), ("valid", C.c_uint8), ("grant", UlGrant)]


class Mib(C.Structure):
    _fields_ = [("found", C.c_uint8), ("nof_ports", C.c_uint8), ("sfn_offset", C.c_uint8), ("phich_length", C.c_uint8), ("phich_resources", C.c_uint8),
                ("bch_payload", C.c_uint8 * 3), ("nof_prb", C.c_uint32), ("sfn", C.c_uint32)]


class UlChest(C.Structure):
    _fields_ = [("noise", C.c_float), ("rsrp", C.c_float), ("snr_db", C.c_float), ("ta_us", C.c_float)]


TAP_UL_SYM = 5
HARQ_NONE, HARQ_NEW, HARQ_RETX = 0, 1, 2            # grant.tb[t].harq_op
HARQ_NEW_TX, HARQ_RE_TX, HARQ_FULL_BUFFER, HARQ_DECODED = 0, 1, 2, 3   # ltephy_harq_classify
HARQ_SLOT_BYTES = 16 * 18448 * 2
UL_FLAG_SLOT1 = 1
CAND_DTYPE = np.dtype([("bits", "<u8"), ("rnti", "<u2"), ("valid", "u1"), ("pad", "u1", 5)])
assert CAND_DTYPE.itemsize == C.sizeof(Cand) == 16
COMPACT_CAP = 248
NEED_FULL_TABLE = -3
COMPACT_DTYPE = np.dtype([("count", "<u4"), ("reserved", "<u4"), ("loc", [("off", "<u2"), ("mask", "u1"), ("pad", "u1")], MAX_LOC), ("pad", "u1", 8),
                          ("list", CAND_DTYPE, COMPACT_CAP)])   # ltephy_compact_t
assert COMPACT_DTYPE.itemsize == 4624

_lib = None


def load_library(build_if_missing=True):
    """Loads libltephy_b200.so; there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing:
        from . import build as _b
        _b.build()
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libltephy_b200.so is missing: build it with __graft_entry__.build(); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    P = C.c_void_p
    L.ltephy_create.argtypes = [C.POINTER(Cfg), C.POINTER(P)]
    L.ltephy_destroy.argtypes = [P]
    L.ltephy_last_error.restype = C.c_char_p
    for f in ("ltephy_sf_len", "ltephy_nof_sizes"):
        getattr(L, f).argtypes = [P]
        getattr(L, f).restype = C.c_uint32
    for f in ("ltephy_nof_cce", "ltephy_dci_size", "ltephy_size_index"):
        getattr(L, f).argtypes = [P, C.c_uint32]
        getattr(L, f).restype = C.c_uint32
    L.ltephy_locations.argtypes = [P, C.c_uint32, P, P, C.c_uint32]
    L.ltephy_locations.restype = C.c_uint32
    L.ltephy_submit_iq.argtypes = [P, P, P, C.c_uint32]
    L.ltephy_submit_iq_device.argtypes = [P, P, P, C.c_uint32]
    L.ltephy_get_phase_a.argtypes = [P, P, P]
    L.ltephy_get_phase_a_compact.argtypes = [P, P, P]
    L.ltephy_phase_a_compact_buffer.argtypes = [P]
    L.ltephy_phase_a_compact_buffer.restype = P
    L.ltephy_copy_phase_a_device.argtypes = [P, P, P]
    L.ltephy_finalize_info.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32]
    L.ltephy_finalize_info.restype = None
    L.ltephy_submit_grants.argtypes = [P, P, C.c_uint32]
    L.ltephy_get_phase_b.argtypes = [P, P, P, C.c_size_t]
    L.ltephy_copy_phase_b_device.argtypes = [P, P, C.c_size_t, P]
    L.ltephy_dci_sweep.argtypes = [P, P, P, C.c_uint32, P]
    L.ltephy_turbo_batch.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, P, P, P]
    L.ltephy_tap.argtypes = [P, C.c_int, P, C.c_size_t]
    L.ltephy_last_timing.argtypes = [P, P]
    L.ltephy_set_ul_cfg.argtypes = [P, C.POINTER(UlCfg)]
    L.ltephy_submit_ul.argtypes = [P, P, P, C.c_uint32, P, C.c_uint32]
    L.ltephy_get_ul.argtypes = [P, P, P, P, C.c_size_t]
    L.ltephy_mark.argtypes = [P, C.c_int]
    L.ltephy_mark_elapsed_ms.argtypes = [P]
    L.ltephy_mark_elapsed_ms.restype = C.c_float
    L.ltephy_last_turbo_work.argtypes = [P, P, P, P]
    L.ltephy_launch_count.argtypes = [P]
    L.ltephy_launch_count.restype = C.c_uint64
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class LtePhy:
    """One PHY context (cell + batch capacity) on one GPU."""

    def __init__(self, nof_prb, nof_ports, cell_id, nof_rx, max_subframes=16, turbo_max_iter=8, device=0, flags=0, symbol_sz=0, phich_resources=0, phich_length=0):
        self.L = load_library()
        cfg = Cfg(nof_prb=nof_prb, nof_ports=nof_ports, cell_id=cell_id, nof_rx=nof_rx, max_subframes=max_subframes,
                  turbo_max_iter=turbo_max_iter, device=device, flags=flags, symbol_sz=symbol_sz, phich_resources=phich_resources, phich_length=phich_length)
        self.cfg = cfg
        self.h = C.c_void_p()
        r = self.L.ltephy_create(C.byref(cfg), C.byref(self.h))
        if r != 0:
            raise RuntimeError("ltephy_create failed (%d): %s" % (r, self.L.ltephy_last_error().decode()))
        self.sf_len = self.L.ltephy_sf_len(self.h)
        self.nsc = 12 * nof_prb
        self.n = 0

    def close(self):
        if self.h:
            self.L.ltephy_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, r, what):
        if r != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, r, self.L.ltephy_last_error().decode()))

    # ---- geometry
    def nof_cce(self, cfi):
        return self.L.ltephy_nof_cce(self.h, cfi)

    def sizes(self):
        return [self.L.ltephy_dci_size(self.h, f) for f in range(NOF_FORMATS)], [self.L.ltephy_size_index(self.h, f) for f in range(NOF_FORMATS)]

    def locations(self, cfi):
        nc = np.zeros(MAX_LOC, np.uint16)
        L = np.zeros(MAX_LOC, np.uint8)
        n = self.L.ltephy_locations(self.h, cfi, _p(nc), _p(L), MAX_LOC)
        return nc[:n].copy(), L[:n].copy()

    # ---- phase A
    def submit_iq(self, iq, tti):
        """iq: complex64 [n][nof_rx][sf_len]"""
        iq = np.ascontiguousarray(iq, np.complex64)
        tti = np.ascontiguousarray(tti, np.uint32)
        n = len(tti)
        assert iq.shape == (n, self.cfg.nof_rx, self.sf_len), iq.shape
        self._keep = (iq, tti)
        self._chk(self.L.ltephy_submit_iq(self.h, _p(iq), _p(tti), n), "submit_iq")
        self.n = n

    def get_phase_a(self, want_cands=True):
        info = (SfInfo * self.n)()
        cands = np.zeros((self.n, MAX_LOC, MAX_SIZES), CAND_DTYPE) if want_cands else None
        self._chk(self.L.ltephy_get_phase_a(self.h, info, _p(cands) if want_cands else None), "get_phase_a")
        return info, cands

    def get_phase_a_compact(self):
        """-> (info, survivor forms as a COMPACT_DTYPE array [n])"""
        info = (SfInfo * self.n)()
        comp = np.zeros(self.n, COMPACT_DTYPE)
        self._chk(self.L.ltephy_get_phase_a_compact(self.h, info, _p(comp)), "get_phase_a_compact")
        return info, comp

    def set_cfo(self, cfo_hz):
        self.L.ltephy_set_cfo.argtypes = [C.c_void_p, C.c_float]
        self._chk(self.L.ltephy_set_cfo(self.h, cfo_hz), "set_cfo")

    def mib_decode(self):
        out = (Mib * self.n)()
        self.L.ltephy_mib_decode.argtypes = [C.c_void_p, C.c_void_p]
        self._chk(self.L.ltephy_mib_decode(self.h, out), "mib_decode")
        return out

    def harq_reserve(self, nslots):
        self.L.ltephy_harq_reserve.argtypes = [C.c_void_p, C.c_uint32]
        self._chk(self.L.ltephy_harq_reserve(self.h, nslots), "harq_reserve")

    def tap(self, what, shape, dtype):
        out = np.zeros(shape, dtype)
        self._chk(self.L.ltephy_tap(self.h, what, _p(out), out.nbytes), "tap")
        return out

    # ---- phase B
    def submit_grants(self, grants):
        arr = (Grant * max(1, len(grants)))(*grants)
        self._grants = arr
        self._ng = len(grants)
        self._chk(self.L.ltephy_submit_grants(self.h, arr, len(grants)), "submit_grants")

    def get_phase_b(self, payload_cap=None):
        res = (TbResult * max(1, 2 * self._ng))()
        cap = payload_cap or (self._ng * 2 * 13000 + 64)
        pl = np.zeros(cap, np.uint8)
        self._chk(self.L.ltephy_get_phase_b(self.h, res, _p(pl), cap), "get_phase_b")
        return res, pl

    # ---- uplink
    def set_ul_cfg(self, n_dmrs1=0, delta_ss=0, group_hopping=0, seq_hopping=0):
        cfg = UlCfg(n_dmrs1=n_dmrs1, delta_ss=delta_ss, group_hopping=group_hopping, seq_hopping=seq_hopping)
        self._chk(self.L.ltephy_set_ul_cfg(self.h, C.byref(cfg)), "set_ul_cfg")

    def decode_ul(self, iq_ul, tti, grants):
        """iq_ul complex64 [n][sf_len] (None: decode the subframes of the previous call again); grants: list of UlGrant -> (results, chest, payload)"""
        iq_ul = None if iq_ul is None else np.ascontiguousarray(iq_ul, np.complex64)
        tti = np.ascontiguousarray(tti, np.uint32)
        arr = (UlGrant * max(1, len(grants)))(*grants)
        self._chk(self.L.ltephy_submit_ul(self.h, None if iq_ul is None else _p(iq_ul), _p(tti), len(tti), arr, len(grants)), "submit_ul")
        res = (TbResult * max(1, len(grants)))()
        ch = (UlChest * max(1, len(grants)))()
        pl = np.zeros(len(grants) * 10000 + 64, np.uint8)
        self._chk(self.L.ltephy_get_ul(self.h, res, ch, _p(pl), pl.nbytes), "get_ul")
        self.n_ul = len(tti)
        return res, ch, pl

    # ---- stand-alone kernels
    def dci_sweep(self, llr, cfi):
        llr = np.ascontiguousarray(llr, np.float32)
        cfi = np.ascontiguousarray(cfi, np.uint32)
        n = len(cfi)
        assert llr.shape == (n, LLR_STRIDE)
        cands = np.zeros((n, MAX_LOC, MAX_SIZES), CAND_DTYPE)
        self._chk(self.L.ltephy_dci_sweep(self.h, _p(llr), _p(cfi), n, _p(cands)), "dci_sweep")
        return cands

    def turbo_batch(self, d, K, max_iter, crc_type):
        d = np.ascontiguousarray(d, np.int16)
        ncb = d.shape[0]
        assert d.shape == (ncb, 3 * (K + 4))
        bits = np.zeros((ncb, K), np.uint8)
        iters = np.zeros(ncb, np.uint8)
        ok = np.zeros(ncb, np.uint8)
        self._chk(self.L.ltephy_turbo_batch(self.h, _p(d), K, ncb, max_iter, crc_type, _p(bits), _p(iters), _p(ok)), "turbo_batch")
        return bits, iters, ok

    def timing(self):
        t = (C.c_float * 4)()
        self.L.ltephy_last_timing(self.h, t)
        return list(t)

    def mark(self, slot):
        self._chk(self.L.ltephy_mark(self.h, slot), "mark")

    def mark_elapsed_ms(self):
        return float(self.L.ltephy_mark_elapsed_ms(self.h))

    def turbo_work(self):
        b, c, i = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self.L.ltephy_last_turbo_work(self.h, C.byref(b), C.byref(c), C.byref(i))
        return b.value, c.value, i.value

    def launch_count(self):
        return int(self.L.ltephy_launch_count(self.h))


def cand_bits(c, nbits):
    """uint64 payload -> array of nbits bits (bit i at position 63-i)."""
    v = int(c)
    return np.array([(v >> (63 - i)) & 1 for i in range(nbits)], np.uint8)


# ------------------------------------------------------------------------------------------------
# host search / grant conversion / one-call pipeline (include/ltephy_search.h)
class DciOut(C.Structure):
    _fields_ = [("sf", C.c_uint32), ("rnti", C.c_uint16), ("format", C.c_uint8), ("L", C.c_uint8), ("ncce", C.c_uint16),
                ("nof_bits", C.c_uint16), ("bits", C.c_uint64), ("histogram_value", C.c_uint32)]


class SearchStats(C.Structure):
    _fields_ = [("nof_decoded_locations", C.c_uint32), ("nof_cce", C.c_uint32), ("nof_missed_cce", C.c_uint32), ("nof_subframes", C.c_uint32),
                ("nof_subframe_collisions_dw", C.c_uint32), ("nof_subframe_collisions_up", C.c_uint32), ("nof_locations", C.c_uint32)]


class DciFields(C.Structure):
    _fields_ = [("rnti", C.c_uint16), ("format", C.c_uint8), ("alloc_type", C.c_uint8), ("mcs", C.c_uint8 * 2), ("rv", C.c_uint8 * 2),
                ("ndi", C.c_uint8 * 2), ("harq_pid", C.c_uint8), ("tpc", C.c_uint8), ("tb_cw_swap", C.c_uint8), ("pinfo", C.c_uint8),
                ("nof_prb", C.c_uint32)]


DCI_DTYPE = np.dtype({"names": ["sf", "rnti", "format", "L", "ncce", "nof_bits", "bits", "histogram_value"],
                      "formats": ["<u4", "<u2", "u1", "u1", "<u2", "<u2", "<u8", "<u4"],
                      "offsets": [DciOut.sf.offset, DciOut.rnti.offset, DciOut.format.offset, DciOut.L.offset, DciOut.ncce.offset,
                                  DciOut.nof_bits.offset, DciOut.bits.offset, DciOut.histogram_value.offset],
                      "itemsize": C.sizeof(DciOut)})
assert DCI_DTYPE.itemsize == C.sizeof(DciOut)
SEQ_NONE = (1 << 64) - 1


def _bind_search(L):
    if getattr(L, "_search_bound", False):
        return
    P = C.c_void_p
    L.ltephy_search_create.argtypes = [P, C.c_uint32]
    L.ltephy_search_create.restype = P
    L.ltephy_search_create_cell.argtypes = [C.c_uint32] * 5
    L.ltephy_search_create_cell.restype = P
    L.ltephy_search_create_cell_ng.argtypes = [C.c_uint32] * 6
    L.ltephy_search_create_cell_ng.restype = P
    L.ltephy_ctrl_region_map.argtypes = [C.c_uint32] * 5 + [P, C.c_uint32, P, P]
    L.ltephy_search_destroy.argtypes = [P]
    L.ltephy_search_config.argtypes = [P, C.c_int, C.c_int, C.c_uint32]
    L.ltephy_search_speculate_256qam.argtypes = [P, C.c_int]
    L.ltephy_search_keep_reserved_mcs.argtypes = [P, C.c_int]
    L.ltephy_search_set_ul_mode.argtypes = [P, C.c_int, C.c_uint16]
    L.ltephy_search_set_ul_hopping.argtypes = [P, C.c_uint32]
    L.ltephy_shard_set_gather_capacity.argtypes = [P, C.c_uint32]
    L.ltephy_search_add_evergreen.argtypes = [P, C.c_uint16, C.c_uint16, C.c_uint32]
    L.ltephy_search_add_forbidden.argtypes = [P, C.c_uint16, C.c_uint16, C.c_uint32]
    L.ltephy_search_activate.argtypes = [P, C.c_uint16, C.c_uint32, C.c_int]
    L.ltephy_search_subframe.argtypes = [P, P, P, C.c_uint32, P, C.c_uint32, P]
    L.ltephy_search_subframe_compact.argtypes = [P, P, P, C.c_uint32, P, C.c_uint32, P]
    L.ltephy_compact_from_table.argtypes = [P, P, P, P]
    L.ltephy_search_batch_compact.argtypes = [P, P, P, P, C.c_uint32, P, C.c_uint32, P]
    L.ltephy_search_needs_full_table.argtypes = [P, P, C.c_uint32]
    L.ltephy_search_get_stats.argtypes = [P, P]
    L.ltephy_search_validate_location.argtypes = [C.c_uint32] * 4 + [C.c_uint16]
    L.ltephy_search_validate_location.restype = C.c_uint32
    L.ltephy_search_rnti_validate_and_refresh.argtypes = [P, C.c_uint16, C.c_uint32]
    L.ltephy_search_rnti_add_candidate.argtypes = [P, C.c_uint16, C.c_uint32]
    L.ltephy_search_rnti_step_time.argtypes = [P]
    for f in ("ltephy_search_rnti_frequency", "ltephy_search_rnti_is_forbidden", "ltephy_search_rnti_is_evergreen"):
        getattr(L, f).argtypes = [P, C.c_uint16, C.c_uint32]
        getattr(L, f).restype = C.c_uint32
    L.ltephy_search_rnti_assoc_format.argtypes = [P, C.c_uint16]
    L.ltephy_search_rnti_assoc_format.restype = C.c_uint32
    L.ltephy_search_rnti_reason.argtypes = [P, C.c_uint16]
    L.ltephy_dci_to_grant.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_int, P, P]
    L.ltephy_ul_dci_to_grant.argtypes = [P, P, C.c_int, P]
    L.ltephy_ul_decode_plan.argtypes = [P, P, C.c_int, P, P]
    L.ltephy_rar_unpack.argtypes = [P, P, C.c_uint32, P, C.c_uint32, P, P]
    L.ltephy_ul_cqi_len.argtypes = [C.c_uint32, C.c_int]
    L.ltephy_ul_uci_layout.argtypes = [P, P, P, P, P]
    L.ltephy_ul_grants_from_dcis.argtypes = [P, P, P, C.c_uint32, P, C.c_uint32, P, P, P, C.c_uint32, P]
    L.ltephy_decode_subframes.argtypes = [P, P, P, P, C.c_uint32, C.c_uint64, P, P, P, C.c_uint32, P, P, P, C.c_size_t]
    L.ltephy_decode_subframes_device.argtypes = [P, P, P, P, C.c_uint32, C.c_uint64, P, P, P, C.c_uint32, P, P, P, C.c_size_t]
    L.ltephy_search_batch.argtypes = [P, P, P, C.c_uint32, P, C.c_uint32, P]
    L.ltephy_grants_from_dcis.argtypes = [P, P, P, C.c_uint32, C.c_uint32, C.c_uint32, P, P, C.c_uint32, P]
    # sharded operation (include/ltephy_shard.h)
    L.ltephy_packed_size.argtypes = [C.c_uint32, C.c_uint32]
    L.ltephy_packed_size.restype = C.c_size_t
    L.ltephy_pack_subframes.argtypes = [P, P, P, C.c_uint32, P, C.c_size_t, P]
    L.ltephy_search_batch_packed.argtypes = [P, P, P, P, C.c_uint32, C.c_uint32, P, C.c_uint32, P, P]
    L.ltephy_packed_needs_full_table.argtypes = [P, P, P, C.c_uint32, C.c_uint32]
    L.ltephy_grants_from_dcis_tc.argtypes = [P, P, P, C.c_uint32, C.c_uint32, C.c_uint32, P, P, C.c_uint32, P]
    L.ltephy_shard_unique_id.argtypes = [P]
    L.ltephy_shard_create.argtypes = [P, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(P)]
    L.ltephy_shard_destroy.argtypes = [P]
    L.ltephy_decode_subframes_sharded.argtypes = [P, P, P, P, C.c_int, P, C.c_uint32, C.c_uint64, P, P, C.c_uint32, P, P, P, C.c_size_t, P]
    L.ltephy_pack_phase_a.argtypes = [P, P, C.c_size_t, P]
    L._search_bound = True


class Search:
    """FALCON acceptance walk + RNTI history (host only; usable without a GPU)."""

    def __init__(self, nof_prb, nof_ports, cell_id, nof_rx, threshold=5, phich_resources=0):
        self.L = load_library()
        _bind_search(self.L)
        self.h = self.L.ltephy_search_create_cell_ng(nof_prb, nof_ports, cell_id, nof_rx, phich_resources, threshold)
        if not self.h:
            raise RuntimeError("ltephy_search_create_cell failed")

    def close(self):
        if self.h:
            self.L.ltephy_search_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def config(self, shortcut=1, skip_secondary=0, update_interval=500):
        self.L.ltephy_search_config(self.h, shortcut, skip_secondary, update_interval)

    def subframe(self, info, cands, sf_in_batch=0, max_out=64):
        """info: SfInfo, cands: CAND_DTYPE array [MAX_LOC][MAX_SIZES] -> accepted DCIs (structured array)"""
        out = np.zeros(max_out, DCI_DTYPE)
        n = C.c_uint32(0)
        cands = np.ascontiguousarray(cands)
        r = self.L.ltephy_search_subframe(self.h, C.byref(info), _p(cands), sf_in_batch, _p(out), max_out, C.byref(n))
        if r < 0:
            raise RuntimeError("ltephy_search_subframe failed (%d)" % r)
        return out[:n.value].copy()

    def compact_from_table(self, info, cands):
        """host restatement of the GPU's survivor selection for one subframe -> COMPACT_DTYPE scalar array [1]"""
        out = np.zeros(1, COMPACT_DTYPE)
        cands = np.ascontiguousarray(cands)
        r = self.L.ltephy_compact_from_table(self.h, C.byref(info), _p(cands), _p(out))
        if r < 0:
            raise RuntimeError("ltephy_compact_from_table failed (%d)" % r)
        return out

    def subframe_compact(self, info, comp, sf_in_batch=0, max_out=64):
        """walk over the survivor form; returns None when the full table is needed (nothing consumed)"""
        out = np.zeros(max_out, DCI_DTYPE)
        n = C.c_uint32(0)
        comp = np.ascontiguousarray(comp)
        r = self.L.ltephy_search_subframe_compact(self.h, C.byref(info), _p(comp), sf_in_batch, _p(out), max_out, C.byref(n))
        if r == NEED_FULL_TABLE:
            return None
        if r < 0:
            raise RuntimeError("ltephy_search_subframe_compact failed (%d)" % r)
        return out[:n.value].copy()

    def stats(self):
        st = SearchStats()
        self.L.ltephy_search_get_stats(self.h, C.byref(st))
        return st

    def dci_to_grant(self, dci_row, sf_idx, cfi, use_256qam=0):
        d = DciOut(sf=int(dci_row["sf"]), rnti=int(dci_row["rnti"]), format=int(dci_row["format"]), L=int(dci_row["L"]), ncce=int(dci_row["ncce"]),
                   nof_bits=int(dci_row["nof_bits"]), bits=int(dci_row["bits"]), histogram_value=int(dci_row["histogram_value"]))
        g = Grant()
        f = DciFields()
        r = self.L.ltephy_dci_to_grant(self.h, C.byref(d), sf_idx, cfi, use_256qam, C.byref(g), C.byref(f))
        return r, g, f


def ul_dci_to_grant(search, dci_row, enable_64qam=1):
    """accepted format-0 DCI -> (return code, UlGrant) via ltephy_ul_dci_to_grant"""
    d = DciOut(sf=int(dci_row["sf"]), rnti=int(dci_row["rnti"]), format=int(dci_row["format"]), L=int(dci_row["L"]), ncce=int(dci_row["ncce"]),
               nof_bits=int(dci_row["nof_bits"]), bits=int(dci_row["bits"]), histogram_value=int(dci_row["histogram_value"]))
    g = UlGrant()
    r = search.L.ltephy_ul_dci_to_grant(search.h, C.byref(d), enable_64qam, C.byref(g))
    return r, g


def ul_decode_plan(search, dci_row, mcs_mod):
    """the decode attempts PUSCH_Decoder::decode makes for this DCI: [(reading, UlGrant)] via ltephy_ul_decode_plan"""
    d = DciOut(sf=int(dci_row["sf"]), rnti=int(dci_row["rnti"]), format=int(dci_row["format"]), L=int(dci_row["L"]), ncce=int(dci_row["ncce"]),
               nof_bits=int(dci_row["nof_bits"]), bits=int(dci_row["bits"]), histogram_value=int(dci_row["histogram_value"]))
    g = (UlGrant * 3)()
    rd = (C.c_uint8 * 3)()
    n = search.L.ltephy_ul_decode_plan(search.h, C.byref(d), mcs_mod, g, rd)
    if n < 0:
        raise ValueError("ltephy_ul_decode_plan: %d" % n)
    return [(int(rd[k]), UlGrant.from_buffer_copy(g[k])) for k in range(n)]


def rar_unpack(search, pdu, max_out=16):
    """MAC RAR PDU -> (return code, [Rar], backoff indicator or -1) via ltephy_rar_unpack"""
    out = (Rar * max_out)()
    n = C.c_uint32(0)
    bo = C.c_int(-1)
    buf = (C.c_uint8 * max(1, len(pdu))).from_buffer_copy(bytes(pdu) if len(pdu) else b"\0")
    r = search.L.ltephy_rar_unpack(search.h, buf, len(pdu), out, max_out, C.byref(n), C.byref(bo))
    return r, [Rar.from_buffer_copy(out[k]) for k in range(n.value)], bo.value


def ul_grants_from_dcis(search, info, dcis, ue_cfgs=()):
    """accepted DCIs of a downlink batch -> [(dci index, reading, UlGrant)] via ltephy_ul_grants_from_dcis (grant.sf = dci.sf + 4)"""
    nd = len(dcis)
    cap = 3 * nd + 1
    g = (UlGrant * cap)()
    gd = np.zeros(cap, np.uint32)
    rd = np.zeros(cap, np.uint8)
    ng = C.c_uint32(0)
    ue = (UlUeCfg * max(1, len(ue_cfgs)))(*ue_cfgs)
    r = search.L.ltephy_ul_grants_from_dcis(search.h, info, dcis.ctypes.data_as(C.c_void_p), nd, ue, len(ue_cfgs), g, gd.ctypes.data_as(C.c_void_p),
                                            rd.ctypes.data_as(C.c_void_p), cap, C.byref(ng))
    if r != 0:
        raise ValueError("ltephy_ul_grants_from_dcis: %d" % r)
    return [(int(gd[k]), int(rd[k]), UlGrant.from_buffer_copy(g[k])) for k in range(ng.value)]


def decode_subframes(phy, search, iq, tti, seq=SEQ_NONE, max_dcis=None, scratch=None):
    """One call through the reference-facing pipeline: host IQ -> (info, dcis, tb results, payload)."""
    L = phy.L
    _bind_search(L)
    iq = np.ascontiguousarray(iq, np.complex64)
    tti = np.ascontiguousarray(tti, np.uint32)
    n = len(tti)
    max_dcis = max_dcis or 32 * n
    if scratch is None:
        scratch = dict(info=(SfInfo * n)(), cands=np.zeros((n, MAX_LOC, MAX_SIZES), CAND_DTYPE), dcis=np.zeros(max_dcis, DCI_DTYPE),
                       tbs=(TbResult * (2 * max_dcis))(), payload=np.zeros(max_dcis * 2 * 2048 + n * 40000, np.uint8))
    nd = C.c_uint32(0)
    r = L.ltephy_decode_subframes(phy.h, search.h, _p(iq), _p(tti), n, seq, scratch["info"], _p(scratch["cands"]), _p(scratch["dcis"]), max_dcis,
                                  C.byref(nd), scratch["tbs"], _p(scratch["payload"]), scratch["payload"].nbytes)
    if r != 0:
        raise RuntimeError("ltephy_decode_subframes failed (%d): %s" % (r, L.ltephy_last_error().decode()))
    phy.n = n
    return scratch["info"], scratch["dcis"][:nd.value], scratch["tbs"], scratch["payload"]


# ------------------------------------------------------------------------------------------------
# sharded operation (include/ltephy_shard.h)
SHARD_ID_BYTES = 384
PACK_MAX_BYTES = 64 + 4 * MAX_LOC + 16 * COMPACT_CAP
PACKED_HDR_DTYPE = np.dtype([("count", "<u4"), ("tti", "<u4"), ("cfi", "<u4"), ("nloc", "<u4"), ("noise", "<f4", (2, 2)), ("rsrp", "<f4", (2, 2)),
                             ("low", "<u8", 2)])
assert PACKED_HDR_DTYPE.itemsize == 64


class ShardStats(C.Structure):
    _fields_ = [("host_ms", C.c_double * 8), ("exchanged_bytes", C.c_uint64), ("n_grants", C.c_uint32), ("used_full_table", C.c_uint32)]


SHARD_HOST_MS = ["submit_a", "wait_a", "exchange_turn", "walk_wait", "walk", "grants", "phase_b", "gather_turn"]


def pack_subframes(search, info, comp):
    """host restatement of the GPU pack kernel: (SfInfo * n), COMPACT_DTYPE[n] -> (uint8 records, uint32 offsets[n + 1])"""
    n = len(comp)
    out = np.zeros(n * PACK_MAX_BYTES, np.uint8)
    offs = np.zeros(n + 1, np.uint32)
    comp = np.ascontiguousarray(comp)
    r = search.L.ltephy_pack_subframes(search.h, info, _p(comp), n, _p(out), out.nbytes, _p(offs))
    if r != 0:
        raise RuntimeError("ltephy_pack_subframes failed (%d)" % r)
    return out[:offs[n]].copy(), offs


def search_batch_packed(search, bufs, offs, n, max_dcis, full=None):
    """walk turn of a sharded batch: bufs[r] uint8 records / offs[r] uint32[n + 1] of every rank ->
    (dcis (sf = global index), tti_cfi uint32 [n * world][2]); None when the full tables are needed and were not given"""
    world = len(bufs)
    bp = (C.c_void_p * world)(*[b.ctypes.data for b in bufs])
    op = (C.c_void_p * world)(*[o.ctypes.data for o in offs])
    fp = (C.c_void_p * world)(*[f.ctypes.data for f in full]) if full is not None else None
    dcis = np.zeros(max_dcis, DCI_DTYPE)
    tc = np.zeros((n * world, 2), np.uint32)
    nd = C.c_uint32(0)
    r = search.L.ltephy_search_batch_packed(search.h, bp, op, fp, world, n, _p(dcis), max_dcis, C.byref(nd), _p(tc))
    if r == NEED_FULL_TABLE:
        return None
    if r != 0:
        raise RuntimeError("ltephy_search_batch_packed failed (%d)" % r)
    return dcis[:nd.value].copy(), tc


class Harq:
    """host bookkeeping of the HARQ mode (ltephy_harq_*, include/ltephy_search.h): slot numbers and NEW_TX / RE_TX / DECODED decisions"""

    def __init__(self, max_rnti=150):
        self.L = load_library()
        L = self.L
        L.ltephy_harq_create.restype = C.c_void_p
        L.ltephy_harq_create.argtypes = [C.c_uint32]
        L.ltephy_harq_destroy.argtypes = [C.c_void_p]
        L.ltephy_harq_classify.argtypes = [C.c_void_p, C.c_uint16, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.ltephy_harq_update.argtypes = [C.c_void_p, C.c_uint16, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_uint32, C.c_int]
        L.ltephy_harq_last_tbs.restype = C.c_int32
        L.ltephy_harq_last_tbs.argtypes = [C.c_void_p, C.c_uint16, C.c_uint32, C.c_uint32]
        self.h = L.ltephy_harq_create(max_rnti)
        assert self.h

    def classify(self, rnti, pid, tb, ndi, tbs, tti):
        slot = C.c_uint32(0)
        r = self.L.ltephy_harq_classify(self.h, rnti, pid, tb, ndi, tbs, tti, C.byref(slot))
        return r, slot.value

    def last_tbs(self, rnti, pid, tb):
        return self.L.ltephy_harq_last_tbs(self.h, rnti, pid, tb)

    def update(self, rnti, pid, tb, ndi, rv, tbs, tti, decoded):
        self.L.ltephy_harq_update(self.h, rnti, pid, tb, ndi, rv, tbs, tti, 1 if decoded else 0)

    def close(self):
        if self.h:
            self.L.ltephy_harq_destroy(self.h)
            self.h = None


def _prefer_torch_nccl():
    """libltephy_b200 binds NCCL with dlopen("libnccl.so.2").  In a Python process that will also import torch, torch's bundled NCCL has to be the
    copy the process holds (an older system libnccl loaded first makes `import torch` fail on missing symbols), so point LTEPHY_NCCL_LIB at it."""
    if os.environ.get("LTEPHY_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["LTEPHY_NCCL_LIB"] = cand
                return
    except Exception:
        pass


class Shard:
    """NCCL communicators + ordered sections of the sharded pipeline (one per process / GPU)."""

    @staticmethod
    def unique_id():
        _prefer_torch_nccl()
        L = load_library()
        _bind_search(L)
        buf = (C.c_uint8 * SHARD_ID_BYTES)()
        if L.ltephy_shard_unique_id(buf) != 0:
            raise RuntimeError("ltephy_shard_unique_id failed: %s" % L.ltephy_last_error().decode())
        return bytes(buf)

    def __init__(self, uid, rank, world, device):
        _prefer_torch_nccl()
        self.L = load_library()
        _bind_search(self.L)
        self.h = C.c_void_p()
        buf = (C.c_uint8 * SHARD_ID_BYTES).from_buffer_copy(uid)
        r = self.L.ltephy_shard_create(buf, rank, world, device, C.byref(self.h))
        if r != 0:
            raise RuntimeError("ltephy_shard_create failed (%d): %s" % (r, self.L.ltephy_last_error().decode()))
        self.rank, self.world = rank, world

    def close(self):
        if self.h:
            self.L.ltephy_shard_destroy(self.h)
            self.h = None


def pack_phase_a(phy):
    """GPU pack kernel over the current batch of phy -> (uint8 records, uint32 offsets[n + 1])"""
    n = phy.n
    out = np.zeros(n * PACK_MAX_BYTES, np.uint8)
    offs = np.zeros(n + 1, np.uint32)
    _bind_search(phy.L)
    phy._chk(phy.L.ltephy_pack_phase_a(phy.h, _p(out), out.nbytes, _p(offs)), "pack_phase_a")
    return out[:offs[n]].copy(), offs
