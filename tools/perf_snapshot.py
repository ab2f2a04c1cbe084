"""Quick timing of phase A / phase B on a cfg-2 style capture with ground-truth grants (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import ltelib
from helpers import make_capture, to_phy_grant, truth_grants
from ltesniffer_b200 import capi

n_unique = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cell = ltelib.Cell(100, 2, 7, 2)
t0 = time.time()
sim, iq, tti, truths, payloads = make_capture(cell, n_unique, seed=2, cfi=3, nof_ues=150, dl_min=8, dl_max=12, tm=3, mcs_min=17, mcs_max=28, snr_db=28.0, full_band=1)
print("generated %d subframes in %.1fs" % (n_unique, time.time() - t0))
N = n_unique * reps
iqb = np.tile(iq, (reps, 1, 1))
ttib = np.tile(tti, reps)
phy = capi.LtePhy(100, 2, 7, 2, max_subframes=N, turbo_max_iter=8, flags=capi.FLAG_SKIP_LOW_POWER)
tg = truth_grants(cell, truths, tti)
grants = []
for r in range(reps):
    grants += [to_phy_grant(sf + r * n_unique, d.rnti, g) for sf, d, g in tg]
for it in range(3):
    t0 = time.time(); phy.submit_iq(iqb, ttib); info, cands = phy.get_phase_a(); t1 = time.time()
    phy.submit_grants(grants); t2 = time.time(); res, pl = phy.get_phase_b(); t3 = time.time()
    tm = phy.timing()
    ok = sum(res[i].crc for i in range(2 * len(grants)))
    ntb = sum(1 for i in range(2 * len(grants)) if res[i].payload_len)
    its = np.mean([res[i].avg_iters for i in range(2 * len(grants)) if res[i].payload_len])
    print("N=%d  phaseA gpu %.2f ms (wall %.1f)  phaseB gpu %.2f ms (turbo %.2f ms, submit wall %.1f ms, get %.1f)  TB ok %d/%d avg iters %.2f  launches %d" % (
        N, tm[0], (t1 - t0) * 1e3, tm[1], tm[2], (t2 - t1) * 1e3, (t3 - t2) * 1e3, ok, ntb, its, phy.launch_count()))
    print("   -> phase A %.0f sf/s, phase B %.0f sf/s, A+B %.0f sf/s (device time)" % (N / tm[0] * 1e3, N / tm[1] * 1e3, N / (tm[0] + tm[1]) * 1e3))
bits = sum(res[i].payload_len * 8 for i in range(2 * len(grants)))
print("   turbo info bits per batch %.1f Mbit -> %.1f Gbit/s (turbo kernels only)" % (bits / 1e6, bits / tm[2] / 1e6))
