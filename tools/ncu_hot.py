#!/usr/bin/env python
"""Print the SASS of an .ncu-rep's kernel with per-instruction executed counts and stall samples
(ncu --page source --csv), optionally only the hottest lines.  usage: ncu_hot.py rep [min_samples]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot_i = tot_s = 0
out = []
for r in rows[2:]:
    try:
        n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    tot_i += n; tot_s += s
    out.append((n, s, r[ix["Source"]].strip()))
print(f"# total warp-instructions {tot_i}, samples {tot_s}")
for n, s, t in out:
    if s >= mins:
        print(f"{n:>10d} {s:>7d}  {t}")
