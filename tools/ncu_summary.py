#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed): key raw metrics, stall
reasons, and a chunked SASS execution profile.  usage: ncu_summary.py <rep> [chunk]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 200
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr = r[0]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'l1tex__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:62s}", [row[i] for row in r[1:]][:4])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, x in enumerate(rows) if x and x[0] == 'Address']
h = rows[hi[0]]
body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
iS = h.index('Source'); iE = h.index('Instructions Executed'); iSm = h.index('# Samples')
tot = {}
for x in body:
    for i, name in enumerate(h):
        if name.startswith('stall_') and 'Not Issued' not in name and i < len(x):
            try: tot[name] = tot.get(name, 0) + float(x[i])
            except ValueError: pass
ts = sum(tot.values()) or 1
print("stalls:", ' '.join(f"{k[6:]}={100*v/ts:.0f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]))
texec = sum(float(x[iE]) for x in body)
print("sass instrs", len(body), "warp-instrs executed", texec)
for c in range(0, len(body), chunk):
    seg = body[c:c + chunk]
    ex = sum(float(x[iE]) for x in seg); sm = sum(float(x[iSm]) for x in seg)
    if ex == 0 and sm == 0: continue
    ops = collections.Counter(re.split(r'[ .]', x[iS].strip().lstrip('@!P0123456789 ').strip())[0] for x in seg)
    print(f"{c:5d} exec={100*ex/texec:5.1f}% samples={sm:6.0f} ({100*sm/ts:4.1f}%)  " + ' '.join(f"{k}:{v}" for k, v in ops.most_common(6)))
