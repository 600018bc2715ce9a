#!/usr/bin/env python
"""Static SASS instruction census per kernel of libsrrg2b.so (cuobjdump -sass).  Usage: python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "srrg2_slam_interfaces_b200/libsrrg2b.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = {}
ops = ["UBLKCP", "SYNCS", "LDGSTS", "REDUX", "FFMA2", "FMUL2", "FADD2", "UTMALDG", "UTMASTG", "ATOMG", "ATOMS", "RED", "LDG"]
cnt = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(2); cnt[cur]["instrs"] += 1
        for o in ops:
            if op == o: cnt[cur][o] += 1
names = list(cnt)
d = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
def short(n):
    n = re.sub(r"\(.*", "", n); n = n.replace("void ", "").replace("s2b::", "").replace("(int)", "")
    return n
print("SASS census of %s (sm_100a), static instruction counts per kernel (cuobjdump -sass; tools/sass_census.py)" % lib)
print("UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, LDGSTS = cp.async, REDUX = warp integer reduce, FFMA2/FMUL2/FADD2 = packed fp32x2 (sm_100)")
print("The kernels of the default product path of an aligner iteration: check_tiles_kernel, nn_kernel, nn_far_kernel, lin_after_search_kernel, icp_solve_kernel.\n")
print("%-64s %7s " % ("kernel", "instrs") + " ".join("%7s" % o for o in ops))
rows = sorted(zip(d, names), key=lambda x: -cnt[x[1]]["instrs"])
for dn, n in rows:
    if "cub" in dn or "thrust" in dn: continue
    print("%-64s %7d " % (short(dn)[:64], cnt[n]["instrs"]) + " ".join("%7d" % cnt[n][o] for o in ops))
