import sys, time, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import bench, ltelib
cell, iq = bench.generate_capture(6, 6)
tti = np.arange(6, dtype=np.uint32)
ltelib.oracle_pipeline(cell, iq[:1], tti[:1])
t=time.perf_counter(); out = ltelib.oracle_pipeline(cell, iq, tti); dt=time.perf_counter()-t
import hashlib
h=hashlib.sha256()
for dcis,tbs,snr,cfi in out:
    for t_ in tbs:
        if t_: 
            for pl in t_[1]: h.update(bytes(pl[:2000]))
print("%.3f s per subframe" % (dt/6), h.hexdigest()[:16])
