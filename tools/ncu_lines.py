#!/usr/bin/env python
"""Per-source-line totals (warp instructions, stall samples) of the first kernel in an .ncu-rep.
Usage: python tools/ncu_lines.py report.ncu-rep [top] [kernel index]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", str(kidx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hi[0]]
body = [r for r in rows[hi[0] + 1:] if len(r) == len(h)]
ie, sm = h.index('Instructions Executed'), h.index('# Samples')
lines = []
for r in body:
    if r[0].isdigit():
        lines.append((int(r[0]), r[1], int(r[ie]) if r[ie].isdigit() else 0, int(r[sm]) if r[sm].isdigit() else 0))
ti = sum(l[2] for l in lines); ts = sum(l[3] for l in lines)
print("total warp instr %d samples %d" % (ti, ts))
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%5d %5.1f%% inst %5.1f%% smp  %s" % (l[0], 100 * l[2] / max(ti, 1), 100 * l[3] / max(ts, 1), l[1].strip()[:110]))
