"""Summarise an ncu launch list (csv) into per-kernel totals / shares, and key metrics of a .ncu-rep."""
import csv, collections, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value'); ui = H.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi: continue
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        if r[ui] in ('usecond', 'us'): v *= 1e3
        elif r[ui] in ('msecond', 'ms'): v *= 1e6
        k = r[ki].split('(')[0]; agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = []
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append("%-26s n=%4d  total %9.3f ms  avg %9.1f us  share %5.1f%%" % (k, v[0], v[1] / 1e6, v[1] / v[0] / 1e3, 100 * v[1] / tot))
    return "\n".join(out)

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'smsp__warp_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_pipe_alu', 'smsp__inst_executed_pipe_fma', 'sm__inst_executed_pipe_lsu',
        'smsp__average_warps_issue_stalled', 'smsp__average_warp_latency_issue_stalled', 'launch__shared_mem_per_block', 'lts__t_bytes.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']

def rep(path):
    r = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    H = rows[0]; units = rows[1]
    out = []
    for row in rows[2:]:
        out.append("== %s  grid %s block %s" % (row[H.index('Kernel Name')][:60], row[H.index('Grid Size')] if 'Grid Size' in H else '', row[H.index('Block Size')] if 'Block Size' in H else ''))
        for i, h in enumerate(H):
            if any(k in h for k in KEYS):
                out.append("   %-78s %s %s" % (h, row[i], units[i]))
    return "\n".join(out)

if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("#", p)
        print(launches(p) if p.endswith('.csv') else rep(p))
