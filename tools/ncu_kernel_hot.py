#!/usr/bin/env python
"""tools/ncu_kernel_hot.py <report.ncu-rep> [top] — headline metrics, stall mix and the hottest source lines / SASS
instructions of one `ncu --set full --import-source on` capture (any kernel)."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: (rows[2][i], rows[1][i]) for i, h in enumerate(rows[0])}
for k in ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__warps_active.avg.per_cycle_active', 'launch__grid_size', 'launch__block_size',
          'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
          'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']:
    if k in d:
        print(f"{k:70s} {d[k][0]} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
c = collections.Counter()
for r in data:
    for h in stalls:
        c[h] += int(r[ix[h]])
print({k: round(v / tot * 100, 1) for k, v in c.most_common(8)})
cum = 0
for k, r in enumerate(data):
    r.append(k)
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:top_n]:
    n = int(r[ix['# Samples']])
    print(f"{100 * n / tot:5.1f}%  #{r[-1]:5d} x{r[ix['Instructions Executed']]:>10s} {r[ix['Source']][:80]:80s}",
          {h[6:]: round(100 * int(r[ix[h]]) / n) for h in stalls if int(r[ix[h]]) > 0.25 * n})
