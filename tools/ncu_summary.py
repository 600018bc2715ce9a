#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, opcode mix, hot SASS regions."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
print("kernel:", rows[2][hdr.index('Kernel Name')] if len(rows) > 2 else '?')
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print("%-70s %s" % (k, [r[i] for r in rows[1:]]))
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
print("stalls (warps per issue-active):")
for k in stall:
    i = hdr.index(k)
    v = rows[2][i]
    try:
        if float(v) > 0.05:
            print("   %-40s %s" % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    except ValueError:
        pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
if hi:
    h = rows[hi[0]]
    body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
    ie, sc, sm = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    tot = sum(int(r[ie]) for r in body if r[ie].isdigit())
    print("sass lines", len(body), "warp instructions", tot)
    ops, smp = collections.Counter(), collections.Counter()
    for r in body:
        if not r[ie].isdigit():
            continue
        t = r[sc].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] += int(r[ie]); smp[op] += int(r[sm]) if r[sm].isdigit() else 0
    ts = sum(smp.values())
    for op, c in ops.most_common(18):
        print("   %-8s %12d %5.1f%% inst  %5.1f%% samples" % (op, c, 100 * c / tot, 100 * smp[op] / max(ts, 1)))
