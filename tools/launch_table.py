"""Per-iteration kernel durations (us) of the last ICP run in an ncu launch list (gpu__time_duration.sum)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
seq = [(r[ki].split('(')[0].replace('void ', '').replace('s2b::', ''), float(r[vi].replace(',', '')) / 1000)
       for r in rows[hi + 1:] if len(r) > vi]
idx = [i for i, (k, _) in enumerate(seq) if 'icp_init' in k]
start = idx[-1] if idx else 0
it = 0; line = []; tot = 0.0
for k, v in seq[start:]:
    line.append("%s=%.1f" % (k[:18], v)); tot += v
    if 'icp_solve' in k:
        print("%2d sum=%6.1f | %s" % (it, tot, ' '.join(line))); line = []; it += 1; tot = 0.0
if line: print(' '.join(line))
