"""Does a large pinned H2D copy overlap with the PHY kernels of another stream?  (diagnostic for the e2e number)"""
import sys, os, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench, ltelib
from ltesniffer_b200 import capi
import ctypes as C

cell, iq_u = bench.generate_capture(50, 8)
B = 1000
sf_len = iq_u.shape[2]
iq_pin = torch.empty((B, cell.nof_rx, sf_len, 2), dtype=torch.float32, pin_memory=True)
v = iq_pin.numpy().view(np.complex64).reshape(B, cell.nof_rx, sf_len)
for r in range(B // 50):
    v[r * 50:(r + 1) * 50] = iq_u
iq_dev = iq_pin.to("cuda")
tti = (np.arange(B) % 50).astype(np.uint32)
phy = capi.LtePhy(cell.nof_prb, cell.nof_ports, cell.cell_id, cell.nof_rx, max_subframes=B, flags=capi.FLAG_SKIP_LOW_POWER)
L = phy.L

def phase_a(n):
    t0 = time.perf_counter()
    for _ in range(n):
        phy._chk(L.ltephy_submit_iq_device(phy.h, C.c_void_p(iq_dev.data_ptr()), tti.ctypes.data_as(C.c_void_p), B), "a")
        phy.n = B
        phy._chk(L.ltephy_get_phase_a_compact(phy.h, (capi.SfInfo * B)(), None), "g")
    return (time.perf_counter() - t0) / n * 1e3

phase_a(3)
print("phase A alone: %.2f ms" % phase_a(10))
dst = torch.empty_like(iq_pin, device="cuda")
s2 = torch.cuda.Stream()
def copy_alone(n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s2):
        for _ in range(n):
            dst.copy_(iq_pin, non_blocking=True)
    s2.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
copy_alone(2)
print("H2D alone: %.2f ms" % copy_alone(5))
stop = [False]; cnt = [0]
def copier():
    torch.cuda.set_device(0)
    with torch.cuda.stream(s2):
        while not stop[0]:
            dst.copy_(iq_pin, non_blocking=True)
            s2.synchronize(); cnt[0] += 1
th = threading.Thread(target=copier); th.start()
time.sleep(0.05)
c0 = cnt[0]; t0 = time.perf_counter()
pa = phase_a(20)
dt = time.perf_counter() - t0; c1 = cnt[0]
stop[0] = True; th.join()
print("phase A with concurrent H2D: %.2f ms; copies meanwhile: %.2f ms each" % (pa, dt / max(1, c1 - c0) * 1e3))
