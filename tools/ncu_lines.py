#!/usr/bin/env python
"""Per-source-line stall samples of an .ncu-rep captured with --import-source on (run where ncu is installed, no GPU needed):
which lines of OUR code the warps of the profiled kernel were sitting on.  usage: ncu_lines.py <rep> [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = None; hdr = None
agg = []      # (samples, file, line, source, top stall)
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    i_s = hdr.index("# Samples")
    try:
        smp = float(r[i_s])
    except ValueError:
        continue
    if smp <= 0:
        continue
    stalls = {h[6:]: float(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and r[i].replace(".", "").isdigit() and float(r[i]) > 0}
    ts = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    agg.append((smp, fname, int(r[0]), r[1].strip()[:110], " ".join(f"{k}={v:.0f}" for k, v in ts)))
tot = sum(a[0] for a in agg) or 1
print("share of warp-stall samples | line (the combined sass,cuda view does not name the file: grid_knn.cu / knn_core.cuh / dev_math.cuh) | top stall reasons | source")
for smp, f, ln, src, st in sorted(agg, key=lambda a: -a[0])[:top]:
    print(f"{100 * smp / tot:5.1f}%  :{ln:<5d} {st:34s} | {src}")
