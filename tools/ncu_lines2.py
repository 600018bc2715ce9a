#!/usr/bin/env python
"""Per-source-line totals of one kernel of an .ncu-rep: warp instructions, average active lanes, stall samples,
aggregated over the SASS rows of each CUDA line.  Usage: python tools/ncu_lines2.py report.ncu-rep [top] [kernel index]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--launch-skip", str(kidx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
h = rows[hi[0]]
ie, te, sm = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
agg = collections.OrderedDict()
for r in rows[hi[0] + 1:]:
    if len(r) != len(h) or not r[0].isdigit():
        continue
    k = int(r[0])
    a = agg.setdefault(k, [r[1], 0, 0, 0, 0])
    a[1] += int(r[ie] or 0); a[2] += int(r[te] or 0); a[3] += int(r[sm] or 0); a[4] += 1
ti = sum(a[1] for a in agg.values()); tt = sum(a[2] for a in agg.values()); ts = sum(a[3] for a in agg.values())
print("total warp instr %d  thread instr %d  lanes %.1f  samples %d" % (ti, tt, tt / max(ti, 1), ts))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5d %5.1f%% inst lanes %4.1f %5.1f%% smp sass %3d  %s" % (k, 100 * a[1] / max(ti, 1), a[2] / max(a[1], 1), 100 * a[3] / max(ts, 1), a[4], a[0].strip()[:96]))
