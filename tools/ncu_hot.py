#!/usr/bin/env python
"""tools/ncu_hot.py <report.ncu-rep> <frames_per_channel> <channels> — per-instruction cycle attribution of the main
(once-per-frame) path of gc_encode_kernel from an ncu --set full --import-source capture (PC sampling)."""
import csv, collections, subprocess, sys, io
rep, fpc, nch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = {h: rows[2][i] for i, h in enumerate(rows[0])}
cyc = float(d['sm__cycles_elapsed.max'])
per = cyc / fpc
print('ms', d['gpu__time_duration.sum'], 'cycles/frame', round(per, 1), 'warp inst', d['smsp__inst_executed.sum'],
      'inst/frame', round(float(d['smsp__inst_executed.sum']) / (fpc * nch), 1))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
frames = fpc * nch
tot = sum(int(r[ix['# Samples']]) for r in data)
cum = 0
with open('/tmp/hot.txt', 'w') as out:
    for k, r in enumerate(data):
        n = int(r[ix['Instructions Executed']]) / frames
        s = int(r[ix['# Samples']]) / tot * per
        if n > 0.5:
            cum += s
            out.write(f"{k:5d} {n:5.2f} {s:6.1f} {cum:7.1f}  {r[ix['Source']].strip()}\n")
print('main path cycles', round(cum, 1))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
c = collections.Counter()
for r in data:
    for h in stalls:
        c[h] += int(r[ix[h]])
print({k: round(v / tot * 100, 1) for k, v in c.most_common(8)})
