#!/usr/bin/env python3
"""per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export:
   tools/ncu_lines.py file.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hi = [i for i, r in enumerate(rows) if len(r) > 5 and r[0] == 'Line No'][0]
hdr = rows[hi]
ii = hdr.index('Instructions Executed'); wi = hdr.index('# Samples'); ti = hdr.index('Thread Instructions Executed')
lines = [r for r in rows[hi + 1:] if len(r) > ii and r[0] != '' and r[ii].isdigit()]
lines = [r for r in lines if r[wi].isdigit() and r[ti].isdigit()]
tot = sum(int(r[ii]) for r in lines); ts = sum(int(r[wi]) for r in lines)
print("total warp instructions %d, samples %d" % (tot, ts))
for r in lines:
    if int(r[ii]) > tot * minp / 100 or int(r[wi]) > ts * minp / 100:
        print("%5s %6.2f%% inst %6.2f%% samp thr/inst %5.1f | %s" % (r[0], 100 * int(r[ii]) / tot, 100 * int(r[wi]) / ts, int(r[ti]) / max(1, int(r[ii])), r[1][:120]))
