#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 LTE PHY decode path.

Workload (BASELINE.json configs[1]): offline DL decode of a synthetic 20 MHz FDD capture, 2 rx antennas,
2 CRS ports, CFI 3, 150 active C-RNTIs, TM3 (DCI 2A, large-delay CDD, two 64QAM codewords), 8-12 DL DCIs per
subframe partitioning all 100 PRBs.  A "step" = one pass of the hot path over one batch of subframes:
OFDM rx -> CRS channel estimate -> PCFICH/PDCCH LLR -> exhaustive DCI Viterbi table -> host FALCON walk ->
PDSCH demap/descramble/rate-dematch/turbo/CRC.  Metric: subframes/s (whole job, all GPUs).

  value : IQ already resident in HBM when the timed region starts (device time, CUDA events on the library stream)
  e2e   : the same metric through the reference-facing C-ABI call (ltephy_decode_subframes) from pinned HOST
          buffers: H2D copy of the IQ and D2H of DCIs / transport blocks inside the timed region
  --impl reference : the CPU path (oracle port; the reference's own arithmetic lives in srsRAN, which cannot be
          built here -- DESIGN.md) on all host cores, bounded sample per step
"""
import argparse
import ctypes as C
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CELL = dict(nof_prb=100, nof_ports=2, cell_id=7, nof_rx=2)
SIM_KW = dict(seed=2, cfi=3, nof_ues=150, dl_min=8, dl_max=12, tm=3, mcs_min=17, mcs_max=28, snr_db=25.0, full_band=1)
WORKLOAD = "cfg2: offline DL 20 MHz FDD, 150 RNTIs, TM3 2x2 64QAM, CFI 3, 8-12 DCI/sf, 100 PRB full band, SNR 25 dB"


def effective_cores():
    """CPUs this process may actually use: affinity mask, capped by the cgroup CPU quota (a container often sees every core of the
    host in os.cpu_count() but is throttled to a few)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(-(-int(q) // int(p)))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, -(-q // p)))
        except Exception:
            pass
    return max(1, n)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_JSON_OUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout: keep the real stdout for it and point fd 1 at stderr, so that library
    chatter (NCCL prints its version on stdout under NCCL_DEBUG=VERSION/WARN) cannot land in front of the line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def emit_json(obj):
    out = claim_stdout()
    out.write(json.dumps(obj) + "\n")
    out.flush()


# ------------------------------------------------------------------------------------------------------
def generate_capture(n_unique, threads):
    """synthetic eNB capture (sim/, input generator -- not measured): n_unique distinct subframes"""
    import ltelib
    cell = ltelib.Cell(CELL["nof_prb"], CELL["nof_ports"], CELL["cell_id"], CELL["nof_rx"])
    sf_len = ltelib.sim().lte_sf_len(cell.nof_prb)
    iq = np.zeros((n_unique, cell.nof_rx, sf_len), np.complex64)

    def work(chunk):
        s = ltelib.Sim(cell=cell, **SIM_KW)
        for i in chunk:
            x, tr, pl = s.subframe(i)
            iq[i] = x
    chunks = [list(range(t, n_unique, threads)) for t in range(threads)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(work, chunks))
    return cell, iq


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.dev, self.rows, self.p = dev, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._rd, daemon=True).start()
        except Exception:
            self.p = None

    def _rd(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm, mx, reasons = [], [], set()
        rows = self.rows[1:] if len(self.rows) >= 3 else self.rows   # the first sample may predate the load (sampling starts just before the warm-up)
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
_CPU_JOB = {}


def _cpu_chunk(b):
    import ltelib
    cell = ltelib.Cell(CELL["nof_prb"], CELL["nof_ports"], CELL["cell_id"], CELL["nof_rx"])
    iq, idx = _CPU_JOB["iq"], _CPU_JOB["idx"]
    if b[1] > b[0]:
        walk = ltelib.OracleWalk(cell)              # one RNTI history per worker, kept across its whole chunk
        for lo in range(b[0], b[1], 8):
            sel = idx[lo:min(lo + 8, b[1])]
            ltelib.oracle_pipeline(cell, iq[sel], sel.astype(np.uint32), walk=walk)
    return b[1] - b[0]


def cpu_pipeline_rate(cell, iq, idx, workers):
    """CPU oracle pipeline (phase A, FALCON walk on the reference's RNTIManager, PDSCH decode) over the subframes iq[idx[i]]
    (tti = idx[i]): `workers` PROCESSES (fork, so the capture is shared copy-on-write) on disjoint contiguous chunks, each with its
    own RNTI history -- the SubframeWorker-style pool of src/src/Phy.cc:29-54.  -> (sf/s, seconds)"""
    import multiprocessing as mp
    import ltelib
    ltelib.walklib()          # build / load before forking
    n = len(idx)
    _CPU_JOB["iq"], _CPU_JOB["idx"] = iq, np.asarray(idx)
    bounds = [(n * t // workers, n * (t + 1) // workers) for t in range(workers)]
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_chunk, [(0, 0)] * workers)       # spin the workers up outside the timed region
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, bounds, chunksize=1)
        dt = time.perf_counter() - t0
    return n / dt, dt


def run_reference(args):
    """--impl reference: CPU path, bounded sample per step, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = effective_cores()
    per_step = int(args.ref_subframes) or 40 * cores
    cell, iq = generate_capture(64, min(cores, 8))
    idx = np.arange(per_step) % len(iq)
    for _ in range(args.warmup):
        cpu_pipeline_rate(cell, iq, idx[:2 * cores], cores)
    dt = 0.0
    for _ in range(args.steps):
        dt += cpu_pipeline_rate(cell, iq, idx, cores)[1]
    v = per_step * args.steps / dt
    out = {"impl": "reference", "metric": "subframes/s", "value": v, "unit": "subframes/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32+int16", "data": "synthetic",
           "config": {"workload": WORKLOAD, "subframes_per_step": per_step},
           "cpu_baseline": {"value": v, "unit": "subframes/s", "cores": cores, "kind": "port",
                            "sample": "%d subframes per step (scalar C port of the srsRAN chain, no SIMD, + the reference's own RNTIManager; srsRAN itself is not buildable here)" % per_step},
           "e2e": {"value": v, "unit": "subframes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_json(out)


# ------------------------------------------------------------------------------------------------------
def bench_turbo_fixed(capi, phy, K, ncb, iters):
    """cfg-4 (BASELINE.json): turbo decoder alone, ncb code blocks of size K (TBS 75 376 -> C = 13, K = 5824), FIXED number of
    iterations (no CRC stop), conditioned int16 inputs resident in HBM when the timed launch starts.  -> information Mbit/s."""
    rng = np.random.default_rng(4)
    d = rng.integers(-40, 41, size=(ncb, 3 * (K + 4)), dtype=np.int16)   # decode time does not depend on the data without early stop
    best = None
    for _ in range(3):
        phy.turbo_batch(d, K, iters, 0)
        ms = phy.timing()[2]
        best = ms if best is None else min(best, ms)
    bits = ncb * K
    return {"K": K, "code_blocks": ncb, "iterations": iters, "kernel_ms": best, "info_mbit_s": bits / best / 1e3,
            "algorithmic_GB_s": ncb * (3 * (K + 4) * 2 + K // 8) / best / 1e6,
            "note": "device time of the turbo launch (CUDA events); inputs copied to HBM before the timed launch; best of 3"}


def bench_dci_sweep(capi, cell, device, n_sf):
    """cfg-3 (BASELINE.json): blind-DCI sweep, full CCE / aggregation-level candidate space of n_sf subframes in one batch, every
    distinct payload size of the 9 formats -> decodes/s of the Viterbi + survivor-selection kernels."""
    rng = np.random.default_rng(3)
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=n_sf, turbo_max_iter=8, device=device)
    try:
        base = (rng.standard_normal((64, capi.LLR_STRIDE)) * 0.3).astype(np.float32)   # sigma 0.3 noise
        base[:, :72 * 26] += np.sign(rng.standard_normal((64, 72 * 26))).astype(np.float32)   # 30 % of the CCEs carry signal-level LLRs
        llr = np.ascontiguousarray(np.tile(base, (n_sf // 64, 1)))
        cfi = np.full(n_sf, 3, np.uint32)
        best = None
        for _ in range(2):
            phy.dci_sweep(llr, cfi)
            ms = phy.timing()[3]
            best = ms if best is None else min(best, ms)
        nloc = len(phy.locations(3)[0])
        nsz = phy.L.ltephy_nof_sizes(phy.h)
        dec = n_sf * nloc * nsz
        return {"subframes": n_sf, "locations": nloc, "payload_sizes": nsz, "decodes": dec, "kernel_ms": best, "decodes_per_s": dec / best * 1e3}
    finally:
        phy.close()


def bench_tm4_256qam(capi, cell, device, threads):
    """cfg-4 variant: 256QAM TM4 two-codeword PDSCH, one UE over all 100 PRB (MCS 27 of the 256QAM table: TBS 97 896 per codeword,
    C = 16, K = 6144), decoded through the batched pipeline; CFI 1, 38 dB."""
    import ltelib
    kw = dict(seed=4, cfi=1, nof_ues=1, dl_min=1, dl_max=1, tm=4, mcs_min=27, mcs_max=27, snr_db=38.0, full_band=1, alt_table=1)
    n_u, n = 16, 256
    s = ltelib.Sim(cell=cell, **kw)
    sf_len = ltelib.sim().lte_sf_len(cell.nof_prb)
    iq = np.zeros((n_u, cell.nof_rx, sf_len), np.complex64)
    for i in range(n_u):
        iq[i] = s.subframe(i)[0]
    iq = np.ascontiguousarray(np.tile(iq, (n // n_u, 1, 1)))
    tti = (np.arange(n) % n_u).astype(np.uint32)
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=n, turbo_max_iter=8, device=device, flags=capi.FLAG_SKIP_LOW_POWER)
    try:
        srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
        srch.L.ltephy_search_speculate_256qam(srch.h, 1)
        best, ok, tot, bits = None, 0, 0, 0
        for _ in range(3):
            t0 = time.perf_counter()
            info, dcis, tbs, payload = capi.decode_subframes(phy, srch, iq, tti)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            ok = sum(1 for i in range(2 * len(dcis)) if tbs[i].crc)
            tot = sum(1 for i in range(2 * len(dcis)) if tbs[i].payload_len)
            bits = sum(8 * tbs[i].payload_len for i in range(2 * len(dcis)) if tbs[i].crc)
        return {"subframes": n, "tb_total": tot, "tb_crc_ok": ok, "max_tb_bits": max([8 * tbs[i].payload_len for i in range(2 * len(dcis))] + [0]),
                "e2e_s": best, "decoded_info_mbit_s": bits / best / 1e6, "turbo_kernel_ms": phy.timing()[2],
                "note": "whole pipeline from host IQ (wall clock, best of 3); 16 distinct subframes tiled"}
    finally:
        phy.close()


def bench_pusch(capi, cell, device):
    """cfg-5 (BASELINE.json) UL leg: PUSCH of DCI-0 grants (L_prb from the valid_prb_ul set, MCS 10-24, full band shared by 6-10 UEs per subframe,
    a third of them with HARQ-ACK / CQI multiplexed), 256 subframes per call through ltephy_submit_ul / ltephy_get_ul from host IQ."""
    import ltelib
    n_u, n = 16, 256
    ucfg = ltelib.UlCfg(n_dmrs1=3, delta_ss=2)
    s = ltelib.Sim(cell=cell, seed=5, snr_db=25.0, nof_ues=1)
    rng = np.random.default_rng(5)
    iq = np.zeros((n_u, s.sf_len), np.complex64)
    per_sf = []
    for i in range(n_u):
        gr = []
        for g in ltelib.make_ul_grants(cell, rng, int(rng.integers(6, 11)), table=1):
            if not 10 <= g.mcs <= 24:
                continue
            if rng.random() < 0.33:
                g.nof_ack, g.I_offset_ack = int(rng.integers(1, 3)), 9
                if rng.random() < 0.5:
                    g.cqi_len, g.I_offset_cqi, g.ri_len, g.I_offset_ri = 30, 6, 1, 5
            gr.append(g)
        iq[i] = ltelib.sim_ul_subframe(s, i, ucfg, gr)[0]
        per_sf.append(gr)
    iq = np.ascontiguousarray(np.tile(iq, (n // n_u, 1)))
    tti = (np.arange(n) % n_u).astype(np.uint32)
    grants = []
    for i in range(n):
        for g in per_sf[i % n_u]:
            grants.append(capi.UlGrant(sf=i, rnti=g.rnti, qm=g.qm, rv=g.rv, L_prb=g.L_prb, n_prb=g.n_prb, n_dmrs2=g.n_dmrs2, tbs=g.tbs, n_prb_slot1=g.n_prb,
                                       flags=capi.UL_FLAG_SLOT1, nof_ack=g.nof_ack, ri_len=g.ri_len, cqi_len=g.cqi_len, I_offset_ack=g.I_offset_ack,
                                       I_offset_ri=g.I_offset_ri, I_offset_cqi=g.I_offset_cqi))
    phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=n, turbo_max_iter=8, device=device)
    try:
        phy.set_ul_cfg(3, 2)
        best, dev_ms, ok, bits = None, None, 0, 0
        for _ in range(3):
            t0 = time.perf_counter()
            res, ch, payload = phy.decode_ul(iq, tti, grants)
            dt = time.perf_counter() - t0
            if best is None or dt < best:
                best, dev_ms = dt, phy.timing()[1]
            ok = sum(1 for r in res[:len(grants)] if r.crc)
            bits = sum(8 * r.payload_len for r in res[:len(grants)] if r.crc)
        return {"subframes": n, "grants": len(grants), "tb_crc_ok": ok, "prb_per_subframe": float(np.mean([sum(g.L_prb for g in gr) for gr in per_sf])),
                "e2e_s": best, "device_ms": dev_ms, "subframes_per_s": n / best, "grants_per_s": len(grants) / best, "decoded_info_mbit_s": bits / best / 1e6,
                "note": "host IQ -> UL OFDM -> PUSCH (DMRS estimate, equalise, IDFT, demap, UCI de-mux) -> rate-dematch -> turbo -> CRC; wall clock, best of 3; "
                        "device_ms = CUDA events around the same work incl. the H2D copy; 16 distinct subframes tiled"}
    finally:
        phy.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1000, help="subframes per step per GPU")
    ap.add_argument("--unique", type=int, default=1000, help="distinct synthetic subframes (tiled to the batch if fewer; cfg-2 asks for 1000 distinct ones)")
    ap.add_argument("--no-sub-records", action="store_true", help="skip the stand-alone cfg-3 / cfg-4 measurements")
    ap.add_argument("--ref-subframes", type=int, default=0, help="subframes per step for --impl reference (0 = 40 per usable core)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="subframes for the cpu_baseline leg (0 = 400 per usable core, about 10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipelines", type=int, default=0, help="PHY handles driven concurrently (host search of batch k overlaps GPU work of batch k+1); "
                    "0 = 6 (measured on one B200: 4 -> 133 k, 6 -> 138 k, 8 -> 134 k subframes/s; sharded batches also wait for two collectives and the replicated walk)")
    args = ap.parse_args()
    claim_stdout()
    faulthandler.enable()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = effective_cores()
    B = args.batch
    t0 = time.time()
    cell, iq_u = generate_capture(min(args.unique, B), min(cores, 16))
    log("[rank %d] generated %d unique subframes in %.1fs" % (rank, len(iq_u), time.time() - t0))
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # before CUDA is touched: the workers are forked
        try:
            ns = args.cpu_sample or 400 * cores
            rate, dt = cpu_pipeline_rate(cell, iq_u, np.arange(ns) % len(iq_u), cores)
            cpu_base = {"value": rate, "unit": "subframes/s", "cores": cores, "kind": "port",
                        "sample": "%d subframes of the same capture in %.1f s, %d worker processes (scalar C port of the srsRAN chain, no SIMD, + the reference's own RNTIManager; srsRAN's AVX decoders are roughly an order of magnitude faster per core)" % (ns, dt, cores)}
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            cpu_base = {"value": None, "unit": "subframes/s", "cores": cores, "kind": "port", "sample": "failed: %r" % (e,)}
        log("[rank 0] cpu_baseline: %r" % (cpu_base,))
    import torch
    import torch.distributed as dist
    from ltesniffer_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus

    reps = (B + len(iq_u) - 1) // len(iq_u)
    sf_len = iq_u.shape[2]
    # pinned host batch (e2e) and a device-resident copy (value)
    iq_pin = torch.empty((B, cell.nof_rx, sf_len, 2), dtype=torch.float32, pin_memory=True)
    iq_np = iq_pin.numpy().view(np.complex64).reshape(B, cell.nof_rx, sf_len)
    for r in range(reps):
        lo, hi = r * len(iq_u), min(B, (r + 1) * len(iq_u))
        iq_np[lo:hi] = iq_u[:hi - lo]
    iq_dev = iq_pin.to("cuda", non_blocking=False)
    # global subframe numbering: subframe g = i * world + rank (round-robin over the GPUs)
    tti = (np.arange(B, dtype=np.uint32) * world + rank).astype(np.uint32)
    tti_local = (np.arange(B, dtype=np.uint32) % len(iq_u)).astype(np.uint32)  # the tti each subframe was generated for

    T = args.pipelines if args.pipelines > 0 else 6
    phys = [capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=B, turbo_max_iter=8, device=local,
                        flags=capi.FLAG_SKIP_LOW_POWER) for _ in range(T)]
    phy = phys[0]
    srch = capi.Search(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx)
    L = phy.L
    capi._bind_search(L)
    max_dcis = 24 * B * world
    tti_c = np.ascontiguousarray(tti_local)

    def p(t):
        return C.c_void_p(t.data_ptr())

    class Scratch:
        def __init__(self):
            self.info = (capi.SfInfo * B)()
            self.cands = torch.empty((B, capi.MAX_LOC, capi.MAX_SIZES, 16), dtype=torch.uint8, pin_memory=True)
            self.comp = torch.empty((B, capi.COMPACT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
            self.dcis = np.zeros(max_dcis, capi.DCI_DTYPE)
            self.tbs = (capi.TbResult * (2 * max_dcis))()
            self.payload = torch.empty(B * 48000, dtype=torch.uint8, pin_memory=True)
            self.nd = C.c_uint32(0)
    scr = [Scratch() for _ in range(T)]
    info, cands, dcis, tbs, payload, nd = scr[0].info, scr[0].cands, scr[0].dcis, scr[0].tbs, scr[0].payload, scr[0].nd
    seq_next = [0]

    def step_on(t, seq, device_resident):
        fn = L.ltephy_decode_subframes_device if device_resident else L.ltephy_decode_subframes
        src = p(iq_dev) if device_resident else p(iq_pin)
        S = scr[t]
        r = fn(phys[t].h, srch.h, src, tti_c.ctypes.data_as(C.c_void_p), B, seq, S.info, p(S.cands), S.dcis.ctypes.data_as(C.c_void_p), max_dcis,
               C.byref(S.nd), S.tbs, p(S.payload), S.payload.numel())
        if r != 0:
            raise RuntimeError("decode_subframes failed: %s" % L.ltephy_last_error().decode())

    def run_steps(nsteps, device_resident):
        """nsteps batches over the T pipelines (thread t takes batches t, t+T, ...; the walk runs in batch order)"""
        base = seq_next[0]
        seq_next[0] += nsteps
        if T == 1:
            for k in range(nsteps):
                step_on(0, base + k, device_resident)
            return
        def worker(t):
            torch.cuda.set_device(local)
            for k in range(t, nsteps, T):
                step_on(t, base + k, device_resident)
        with ThreadPoolExecutor(T) as ex:
            list(ex.map(worker, range(T)))

    def step_single(device_resident):
        run_steps(1, device_resident)

    # ---- sharded operation (N > 1): subframe g -> GPU g mod N, through the library's own entry point
    # ltephy_decode_subframes_sharded (include/ltephy_shard.h): phase A local; packed survivor forms all-gathered over NCCL; the
    # walk replayed on every rank; phase B local; one NCCL gather of the decoded transport blocks to rank 0.  T host threads,
    # one PHY handle each (thread t takes batches t, t+T, ...); the library orders the exchange / walk / gather sections by seq.
    TB_DTYPE = np.dtype({"names": ["crc", "avg_iters", "nof_cb", "payload_off", "payload_len"], "formats": ["u1", "u1", "<u2", "<u4", "<u4"],
                         "offsets": [0, 1, 2, 4, 8], "itemsize": C.sizeof(capi.TbResult)})
    sh = None
    sh_seq = [0]
    sh_stats = []          # (seq, ShardStats) of every batch of the last run_sharded call
    if world > 1:
        uid = [capi.Shard.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        sh = capi.Shard(uid[0], rank, world, local)
        sh_payload = [torch.empty(B * 24000 * (world if rank == 0 else 1) + (1 << 20), dtype=torch.uint8, pin_memory=True) for _ in range(T)]

    def run_sharded(nsteps, device_resident, T=T):
        base = sh_seq[0]
        sh_seq[0] += nsteps
        del sh_stats[:]
        errs = []
        src = p(iq_dev) if device_resident else p(iq_pin)

        def worker(t):
            torch.cuda.set_device(local)
            for k in range(t, nsteps, T):
                S = scr[t]
                st = capi.ShardStats()
                r = L.ltephy_decode_subframes_sharded(sh.h, phys[t].h, srch.h, src, 1 if device_resident else 0, tti_c.ctypes.data_as(C.c_void_p), B,
                                                      base + k, S.info, S.dcis.ctypes.data_as(C.c_void_p), max_dcis, C.byref(S.nd), S.tbs,
                                                      p(sh_payload[t]), sh_payload[t].numel(), C.byref(st))
                if r != 0:
                    errs.append(RuntimeError("decode_subframes_sharded failed: %s" % L.ltephy_last_error().decode()))
                    return
                sh_stats.append((base + k, t, st))
        if T == 1 or nsteps == 1:
            worker(0)
        else:
            with ThreadPoolExecutor(T) as ex:
                list(ex.map(worker, range(T)))
        if errs:
            raise errs[0]

    def step_sharded(device_resident):
        run_sharded(1, device_resident)

    step = step_sharded if world > 1 else step_single

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi clocks / throttle reasons are sampled every 100 ms from here to the end of the device-resident timed region: the
    # warm-up steps are the same load, and the timed region alone (~0.1 s) would yield a single sample
    clk = ClockSampler(local)
    clk.start()
    log("[rank %d] warm-up" % rank)
    # ---------------- warm-up ----------------
    for _ in range(max(3, args.warmup)):
        if world > 1:
            run_sharded(T, True)
        else:
            run_steps(T, True)
    barrier()
    launches0 = sum(ph.launch_count() for ph in phys)
    log("[rank %d] timed region (device-resident IQ)" % rank)
    # ---------------- value: IQ resident in HBM ----------------
    turbo_ms, phase_a_ms, phase_b_ms = [], [], []
    host_ms = np.zeros(8)
    L.ltephy_last_host_timing.argtypes = [C.c_void_p]
    for ph in phys:
        ph.mark(0)
    t0 = time.perf_counter()
    if world > 1:
        run_sharded(args.steps, True)
    else:
        run_steps(args.steps, True)
    for ph in phys:
        ph.mark(1)
    dev_ms = max(ph.mark_elapsed_ms() for ph in phys)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = sum(ph.launch_count() for ph in phys) - launches0
    for ph in phys:
        tm = ph.timing()
        phase_a_ms.append(tm[0]), phase_b_ms.append(tm[1]), turbo_ms.append(tm[2])
    hm = np.zeros(8)
    L.ltephy_last_host_timing(hm.ctypes.data_as(C.c_void_p))
    host_ms += hm
    clocks = clk.stop()
    t_val = torch.tensor([dev_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t_val, op=dist.ReduceOp.MAX)
    dev_ms = float(t_val.item())
    def tb_counts(S):
        """CRC-ok / total transport blocks of the last batch decoded into scratch S; at N > 1 rank 0 holds the gathered blocks of
        every rank, and (all ranks decode the same capture) must see exactly N times what its own subframes yield"""
        ndv = S.nd.value
        tbv = np.frombuffer(S.tbs, dtype=TB_DTYPE, count=2 * ndv)
        ok = np.repeat(S.dcis["sf"][:ndv] % world == 0, 2)
        return int((tbv["crc"] > 0).sum()), int((tbv["payload_len"] > 0).sum()), int(((tbv["crc"] > 0) & ok).sum())

    def shard_host_ms():
        if not sh_stats:
            return None
        m = np.mean([[st.host_ms[i] for i in range(8)] for _, _, st in sh_stats], axis=0)
        d = dict(zip(capi.SHARD_HOST_MS, [round(float(x), 3) for x in m]))
        d["exchanged_bytes_per_step"] = int(np.mean([st.exchanged_bytes for _, _, st in sh_stats]))
        d["full_table_fallbacks"] = int(sum(st.used_full_table for _, _, st in sh_stats))
        return d
    tb_ok, ntb, tb_ok_own = tb_counts(scr[0])
    sh_host_value = shard_host_ms() if world > 1 else None
    tbytes, ncb, info_bits = phy.turbo_work()
    log("[rank %d] timed region (host IQ)" % rank)
    # ---------------- e2e: host IQ through the C-ABI ----------------
    T_e2e = min(T, 4)   # with the samples coming from the host every batch starts with a 9 ms copy: more than 4 batches in flight only queue behind it
    if world > 1:
        run_sharded(T_e2e, False, T_e2e)
    else:
        run_steps(2 * T, False)
    barrier()
    t0 = time.perf_counter()
    if world > 1:
        run_sharded(args.steps, False, T_e2e)
    else:
        run_steps(args.steps, False)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t_e2e = torch.tensor([e2e_ms], device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_e2e.item())
    e2e_pa = float(np.mean([ph.timing()[0] for ph in phys]))   # H2D copy + phase A of the last e2e batch of each pipeline (CUDA events)
    e2e_pb = float(np.mean([ph.timing()[1] for ph in phys]))
    hm2 = np.zeros(8)
    L.ltephy_last_host_timing(hm2.ctypes.data_as(C.c_void_p))
    e2e_tb_ok = tb_counts(scr[0])[0]    # same count as the device-resident run
    sh_host_e2e = shard_host_ms() if world > 1 else None
    d2h = B * (C.sizeof(capi.SfInfo) + capi.COMPACT_DTYPE.itemsize) + int(info_bits // 8) + 12 * 2 * 24 * B // 8

    # ---------------- sub-records: BASELINE.json configs 3 and 4 through the stand-alone batched entry points (N = 1 only) --------
    sub_records = None
    if world == 1 and rank == 0 and not args.no_sub_records:
        sub_records = {}
        log("[rank 0] sub-records")
        try:
            sub_records["cfg4_turbo_fixed8"] = bench_turbo_fixed(capi, phys[0], 5824, 13 * 2000, 8)
        except Exception as e:
            sub_records["cfg4_turbo_fixed8"] = {"error": repr(e)}
        try:
            sub_records["cfg3_dci_sweep"] = bench_dci_sweep(capi, cell, local, 8192)
        except Exception as e:
            sub_records["cfg3_dci_sweep"] = {"error": repr(e)}
        try:
            sub_records["cfg4_tm4_256qam_2cw"] = bench_tm4_256qam(capi, cell, local, min(cores, 16))
        except Exception as e:
            sub_records["cfg4_tm4_256qam_2cw"] = {"error": repr(e)}
        try:
            sub_records["cfg5_pusch"] = bench_pusch(capi, cell, local)
        except Exception as e:
            sub_records["cfg5_pusch"] = {"error": repr(e)}
        log("[rank 0] sub-records: %r" % (sub_records,))

    # ---------------- roofline of the dominant kernel (turbo decoder) ----------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    t_turbo = float(np.mean(turbo_ms)) * 1e-3
    achieved = tbytes / t_turbo / 1e9 if t_turbo > 0 else 0.0
    traffic = None   # dram__bytes_read + dram__bytes_write of the step's turbo launches, from the committed ncu --set full capture of this workload
    try:
        tt = json.load(open(os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")))["turbo_kernel_step_total"]
        if tt["subframes_per_step"] == B:
            traffic = tt["traffic_bytes"]
    except Exception:
        pass
    # the same launches with nothing else on the GPU (one pipeline, one extra step after the timed regions): with several pipelines
    # the in-region duration above includes the time the turbo CTAs share the SMs with the other streams' kernels
    alone = None
    if world == 1:
        barrier()
        step_on(0, capi.SEQ_NONE, True)
        barrier()
        t_alone = phys[0].timing()[2] * 1e-3
        if t_alone > 0:
            alone = {"launch_ms": t_alone * 1e3, "achieved": tbytes / t_alone / 1e9, "frac": tbytes / t_alone / 1e9 / peak,
                     "turbo_info_mbit_s": info_bits / t_alone / 1e6}
    roofline = {"kernel": "turbo_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "timed_alone": alone,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6.65 TB/s", "launch_ms": t_turbo * 1e3,
                "algorithmic_bytes_per_launch": tbytes, "code_blocks": ncb, "turbo_info_mbit_s": info_bits / t_turbo / 1e6 if t_turbo > 0 else 0.0,
                "note": "max-log-MAP is ALU/issue bound (~1e3 int ops per info bit); the HBM fraction is small by construction (SURVEY.md 8d)"}

    out = None
    if rank == 0:
        value = B * world * args.steps / (dev_ms * 1e-3)
        out = {"metric": "subframes/s", "value": value, "unit": "subframes/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
               "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+int16",
               "data": "synthetic",
               "config": {"workload": WORKLOAD, "subframes_per_step_per_gpu": B, "pipelines": T, "pipelines_e2e": (min(T, 4) if world > 1 else T), "unique_subframes": len(iq_u), "turbo_max_iter": 8,
                          "cache_note": "inputs larger than L2: %.0f MB of IQ per step per GPU" % (iq_pin.numel() * 4 / 1e6),
                          "sharding": "ltephy_decode_subframes_sharded: subframe g -> GPU g mod N; NCCL all-gather of the packed survivor forms; walk replayed on every rank; one NCCL gather of the decoded TBs to rank 0" if world > 1 else "single GPU",
                          "tb_crc_ok": tb_ok, "tb_total": ntb, "tb_crc_ok_own_subframes": tb_ok_own,
                          "tb_check": "gathered CRC-ok count == N x rank 0's own" if tb_ok == world * tb_ok_own else "MISMATCH: %d != %d x %d" % (tb_ok, world, tb_ok_own),
                          "dcis_per_step": int(nd.value)},
               "wall_ms_per_step": wall_ms / args.steps, "phase_a_ms": float(np.mean(phase_a_ms)), "phase_b_ms": float(np.mean(phase_b_ms)),
               "host_ms": sh_host_value if world > 1 else dict(zip(["submit_a", "wait_a", "search", "grants", "submit_b", "wait_b"], [round(float(x), 3) for x in host_ms[:6]])),
               "e2e": {"value": B * world * args.steps / (e2e_ms * 1e-3), "unit": "subframes/s",
                       "h2d_bytes_per_step": int(iq_pin.numel() * 4 + B * 4), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": e2e_ms / args.steps, "tb_crc_ok": e2e_tb_ok, "h2d_plus_phase_a_ms": e2e_pa, "phase_b_ms": e2e_pb,
                       "host_ms": sh_host_e2e if world > 1 else dict(zip(["submit_a", "wait_a", "search", "grants", "submit_b", "wait_b"], [round(float(x), 3) for x in hm2[:6]]))},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "sub_records": sub_records}

    if rank == 0 and cpu_base is not None:
        out["cpu_baseline"] = cpu_base
    if rank == 0:
        emit_json(out)
    if sh is not None:
        sh.close()
    for ph in phys:
        ph.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
