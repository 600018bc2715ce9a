"""CPU model of nn_tile_body's staging: per 256-query tile of the Hilbert-ordered moving cloud, the
dilated cell box (rows, table entries, staged points) -- to size the shared-memory caps.
Usage: python tools/tile_stats.py [n] [tile]"""
import sys
sys.path.insert(0, '.')
import numpy as np
from srrg2_slam_interfaces_b200 import synthetic as syn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
TQ = int(sys.argv[2]) if len(sys.argv) > 2 else 256
order = sys.argv[3] if len(sys.argv) > 3 else "hilbert"
d = syn.make_icp3d(n, n, seed=2)
F, M = d["fixed"], d["moving"]
md, R = 0.3, 2
cell = np.float32(md * 1.05 / R)

def hilbert_keys(q, B=10):
    X = [q[:, 0].astype(np.uint32).copy(), q[:, 1].astype(np.uint32).copy(), q[:, 2].astype(np.uint32).copy()]
    Mb = 1 << (B - 1)
    Q = Mb
    while Q > 1:
        P = np.uint32(Q - 1)
        for i in range(3):
            hit = (X[i] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            t = np.where(hit, 0, (X[0] ^ X[i]) & P).astype(np.uint32)
            X[0] ^= t; X[i] ^= t
        Q >>= 1
    for i in range(1, 3): X[i] ^= X[i - 1]
    t = np.zeros_like(X[0]); Q = Mb
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ np.uint32(Q - 1), t)
        Q >>= 1
    X = [x ^ t for x in X]
    key = np.zeros(len(q), dtype=np.uint64)
    for b in range(B - 1, -1, -1):
        for i in range(3):
            key = (key << np.uint64(1)) | ((X[i] >> np.uint32(b)) & 1).astype(np.uint64)
    return key

def morton_keys(q, B=10):
    key = np.zeros(len(q), dtype=np.uint64)
    for b in range(B - 1, -1, -1):
        for i in (2, 1, 0):
            key = (key << np.uint64(1)) | ((q[:, i].astype(np.uint32) >> np.uint32(b)) & 1).astype(np.uint64)
    return key

mn, mx = M.min(0), M.max(0)
q = np.clip((M - mn) * (1023.0 / (mx - mn)), 0, 1023).astype(np.uint32)
key = hilbert_keys(q) if order == "hilbert" else morton_keys(q)
perm = np.argsort(key, kind="stable")
Ms = M[perm]
fo = F.min(0)
dims = (np.floor((F.max(0) - fo) / cell) + 1).astype(int)
fc = np.clip(np.floor((F - fo) / cell).astype(int), 0, dims - 1)
occ = np.zeros(dims, dtype=np.int32)
np.add.at(occ, (fc[:, 0], fc[:, 1], fc[:, 2]), 1)
# 3D prefix sums for box counts
ps = np.zeros(dims + 1, dtype=np.int64)
ps[1:, 1:, 1:] = occ.cumsum(0).cumsum(1).cumsum(2)
def box_count(lo, hi):
    x0, y0, z0 = lo; x1, y1, z1 = hi + 1
    return (ps[x1, y1, z1] - ps[x0, y1, z1] - ps[x1, y0, z1] - ps[x1, y1, z0]
            + ps[x0, y0, z1] + ps[x0, y1, z0] + ps[x1, y0, z0] - ps[x0, y0, z0])
mc = np.floor((Ms - fo) / cell).astype(int)   # identity transform (iteration 0)
mc = np.clip(mc, -2, dims + 1)
nt = (n + TQ - 1) // TQ
rows, ents, pts = np.zeros(nt, int), np.zeros(nt, int), np.zeros(nt, int)
for t in range(nt):
    c = mc[t * TQ:(t + 1) * TQ]
    lo = np.clip(c.min(0) - R, 0, dims - 1); hi = np.clip(c.max(0) + R, 0, dims - 1)
    b = hi - lo + 1
    rows[t] = b[1] * b[2]; ents[t] = rows[t] * (b[0] + 1); pts[t] = box_count(lo, hi)
for name, v in (("rows", rows), ("entries", ents), ("points", pts)):
    print("%-8s mean %8.0f  p50 %6d  p90 %6d  p99 %6d  max %7d" % (name, v.mean(), *np.percentile(v, [50, 90, 99]).astype(int), v.max()))
for cr, ce, cp in ((768, 5632, 2560), (1024, 8192, 3072), (1536, 12288, 4096), (2048, 16384, 6144)):
    ok = (rows <= cr) & (ents <= ce) & (pts <= cp)
    print("caps rows %d entries %d pts %d: staged %.1f%%  (rows fail %.1f%% entries fail %.1f%% pts fail %.1f%%)" % (
        cr, ce, cp, 100 * ok.mean(), 100 * (rows > cr).mean(), 100 * (ents > ce).mean(), 100 * (pts > cp).mean()))
