"""The C++ host side above the C ABI (include/srrg2b.hpp: CorrespondenceFinderB200, MultiAlignerB200,
PoseGraphSolverB200 mirror the reference's CorrespondenceFinder_ / MultiAlignerBase_ / Solver surface).
CPU: the header compiles as C++17, links against libsrrg2b.so, and the program fails loudly without a
GPU.  GPU: the finder and the aligner driven through those classes give the oracle's results bit for bit."""
import os
import struct
import subprocess

import numpy as np
import pytest

from srrg2_slam_interfaces_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "srrg2_slam_interfaces_b200")


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build_cuda()
    exe = str(tmp_path_factory.mktemp("cpp") / "host_mirror_main")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_main.cpp"), "-o", exe,
                           "-L", LIBDIR, "-lsrrg2b", "-Wl,-rpath," + LIBDIR])
    return exe


def _write_input(path, d):
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", d["fixed"].shape[0], d["moving"].shape[0]))
        for k in ("fixed", "fixed_normals", "moving", "moving_normals"):
            f.write(np.ascontiguousarray(d[k], dtype=np.float32).tobytes())


def test_compiles_and_fails_loudly_without_a_gpu(binary, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    d = syn.make_icp3d(200, 150, seed=3)
    _write_input(tmp_path / "in.bin", d)
    r = subprocess.run([binary, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle(binary, tmp_path, oracle):
    O = oracle
    d = syn.make_icp3d(12000, 9001, seed=21)
    _write_input(tmp_path / "in.bin", d)
    r = subprocess.run([binary, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    buf = open(tmp_path / "out.bin", "rb").read()
    corr_t = np.dtype([("f", "<i4"), ("m", "<i4"), ("r", "<f4")])
    stats_t = np.dtype([("iteration", "<i4"), ("solver_status", "<i4"), ("num_inliers", "<i8"), ("num_outliers", "<i8"),
                        ("num_suppressed", "<i8"), ("num_correspondences", "<i8"), ("chi_inliers", "<f8"),
                        ("chi_outliers", "<f8"), ("num_saturated", "<i8")])
    off = 0
    nf = struct.unpack_from("<q", buf, off)[0]; off += 8
    find = np.frombuffer(buf, corr_t, nf, off); off += nf * corr_t.itemsize
    T = np.frombuffer(buf, "<f4", 16, off).reshape(4, 4); off += 64
    status, ns = struct.unpack_from("<ii", buf, off); off += 8
    stats = np.frombuffer(buf, stats_t, ns, off); off += ns * stats_t.itemsize
    nc = struct.unpack_from("<q", buf, off)[0]; off += 8
    corr = np.frombuffer(buf, corr_t, nc, off); off += nc * corr_t.itemsize
    nr = struct.unpack_from("<i", buf, off)[0]; off += 4
    res_t = np.dtype([("verdict", "<i4"), ("status", "<i4"), ("ncorr", "<i8"), ("ninl", "<i8"), ("chi", "<f4"), ("T", "<f4", (16,))])
    res = np.frombuffer(buf, res_t, nr, off); off += nr * res_t.itemsize
    nd = struct.unpack_from("<i", buf, off)[0]; off += 4
    det_t = np.dtype([("target", "<i4"), ("T", "<f4", (16,))])
    det = np.frombuffer(buf, det_t, nd, off)
    # oracle, same inputs
    F = O.CloudRef(d["fixed"], d["fixed_normals"])
    M = O.CloudRef(d["moving"], d["moving_normals"])
    fp = O.finder_params(0.3, 0.8)
    ix = O.Index(F, O.NN_KDTREE)
    fidx, resp = O.find(ix, F, M, np.eye(4, dtype=np.float32), fp)
    hit = np.nonzero(fidx >= 0)[0]
    assert np.array_equal(find["m"], hit) and np.array_equal(find["f"], fidx[hit]) and np.array_equal(find["r"], resp[hit])
    o = O.icp_run(3, [O.make_slice(F, M, None, fp, O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, 0.01))],
                  O.aligner_params(max_iterations=8, min_num_inliers=10), np.eye(4))
    assert status == o["status"] and ns == len(o["stats"])
    assert np.array_equal(T, np.asarray(o["T"], dtype=np.float32))
    for got, want in zip(stats, o["stats"]):
        for key in stats_t.names:
            assert got[key] == want[key], key
    ofi, omi, ors = o["correspondences"][0]
    assert np.array_equal(corr["f"], ofi) and np.array_equal(corr["m"], omi) and np.array_equal(corr["r"], ors)
    # the loop detector's candidate loop (N3): hints 0 and 2 are live
    half = O.CloudRef(d["moving"][: d["moving"].shape[0] // 2], d["moving_normals"][: d["moving"].shape[0] // 2])
    fa = O.factor_params(O.FACTOR_PLANE, O.ROB_HUBER, 0.01)
    ol = O.closure_loop(3, [[O.make_slice(F, M, None, fp, fa)], [O.make_slice(F, half, None, fp, fa)]],
                        O.aligner_params(max_iterations=8, min_num_inliers=10, enable_inlier_only_runs=True),
                        [np.eye(4, dtype=np.float32)] * 2, 100, 0.005, 0.5)
    assert nr == 2
    for got, want in zip(res, ol):
        assert got["verdict"] == want["verdict"] and got["status"] == want["aligner_status"]
        assert np.array_equal(got["T"].reshape(4, 4), want["T"])
        if want["aligner_status"] == 0:
            assert got["ncorr"] == want["num_correspondences"] and got["ninl"] == want["num_inliers"]
            assert np.float32(got["chi"]) == np.float32(want["chi_inliers"])
    accepted = [(0, 2)[k] for k, want in enumerate(ol) if want["verdict"] == O.CLOSURE_ACCEPT]
    assert list(det["target"]) == accepted and len(accepted) >= 1
    for dd in det:
        assert np.array_equal(dd["T"].reshape(4, 4), ol[(0, 2).index(dd["target"])]["T"])
