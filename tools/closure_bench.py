"""N3: K candidate alignments batched (srrg2b_closure_batch, one context per candidate, fixed side lent) against the
reference's flow -- one aligner, setMoving + compute per candidate, serially.  Wall-clock per batch, uploads included
in neither (clouds resident).  Usage: python tools/closure_bench.py [K] [n_fixed] [n_moving] [dim]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from srrg2_slam_interfaces_b200 import capi as A, synthetic as syn

K = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
nm = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
dim = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rng = np.random.default_rng(3)
cands = []
for k in range(K):
    if dim == 2:
        Ts = syn.iso2(*rng.uniform(-0.1, 0.1, size=2), rng.uniform(-0.03, 0.03))
        d = syn.make_icp2d(nf, nm + 16 * k, seed=9, T_star=Ts, paired=False)
    else:
        Ts = syn.iso3(rng.uniform(-0.08, 0.08, size=3), rng.uniform(-0.02, 0.02, size=3))
        d = syn.make_icp3d(nf, nm + 16 * k, seed=9, T_star=Ts, moving_stream=k + 1)
    cands.append((d, Ts.astype(np.float32)))
fixed, fixed_n = cands[0][0]["fixed"], cands[0][0]["fixed_normals"]
sl = [A.make_slice(dim, 0, None, A.finder_params(0.5 if dim == 2 else 0.3, 0.7), A.factor_params(A.FACTOR_PLANE, A.ROB_CAUCHY, 0.05))]
ap = A.aligner_params(max_iterations=15, min_num_inliers=20)
src = A.Context(dim)
src.set_cloud(A.FIXED, 0, fixed, fixed_n)
src.set_cloud(A.MOVING, 0, cands[0][0]["moving"], cands[0][0]["moving_normals"])
src.icp_run(sl, ap, np.eye(dim + 1))
ctxs = []
for d, Ts in cands:
    x = A.Context(dim)
    x.share_fixed(0, src, 0)
    x.set_cloud(A.MOVING, 0, d["moving"], d["moving_normals"])
    ctxs.append(x)
guesses = [Ts for _, Ts in cands]
cp = A.closure_params(400, 0.01, 0.6)
for rep in range(3):
    res = A.closure_batch(ctxs, sl, ap, guesses, cp)
t = []
for rep in range(10):
    t0 = time.perf_counter(); res = A.closure_batch(ctxs, sl, ap, guesses, cp); t.append(time.perf_counter() - t0)
batch = min(t)
# serial on resident clouds: the same K contexts, one compute() after the other (each waits for its result)
t = []
for rep in range(10):
    t0 = time.perf_counter()
    for x, g in zip(ctxs, guesses):
        x.icp_run(sl, ap, g)
    t.append(time.perf_counter() - t0)
serial = min(t)
# the reference's flow: ONE aligner, setMoving (upload + Hilbert sort) + compute per candidate
t = []
for rep in range(5):
    t0 = time.perf_counter()
    for (d, Ts) in cands:
        src.set_cloud(A.MOVING, 0, d["moving"], d["moving_normals"])
        src.icp_run(sl, ap, Ts)
    t.append(time.perf_counter() - t0)
ref_flow = min(t)
dev = sum(r["device_ms"] for r in res)
print("K=%d dim=%d fixed=%d moving~%d: batch %.3f ms | serial resident %.3f ms | serial setMoving+compute %.3f ms | sum of device ms %.3f | accepted %d"
      % (K, dim, nf, nm, batch * 1e3, serial * 1e3, ref_flow * 1e3, dev, sum(r["verdict"] == 0 for r in res)))
