"""The reference-side adapter classes (adapters/*.h: CorrespondenceFinderB200_, MultiAlignerB200_, PoseGraphSolverB200_)
are real headers: they must instantiate against the stand-in headers of SURVEY.md Appendix A (adapters/stubs) and the
C ABI of include/srrg2b.h.  (With srrg2_core / srrg2_solver installed the same headers build against the real ones.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("unit", ["check_adapters.cpp", "instances_b200.cpp"])
def test_adapters_instantiate_against_the_stub_surface(unit):
    """check_adapters.cpp: explicit instantiations; instances_b200.cpp: the BOSS registration unit (pattern R/instances.cpp:21-85)."""
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                          "-I", os.path.join(ROOT, "adapters", "stubs"), "-I", os.path.join(ROOT, "adapters"),
                          os.path.join(ROOT, "adapters", unit)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]


def test_adapters_name_every_entry_point_they_bind():
    """Every srrg2b_* call in the adapters is declared in include/srrg2b.h."""
    import re
    header = open(os.path.join(ROOT, "include", "srrg2b.h")).read()
    declared = set(re.findall(r"\b(srrg2b_[a-z_0-9]+)\s*\(", header))
    used = set()
    for name in os.listdir(os.path.join(ROOT, "adapters")):
        if name.endswith(".h"):
            used |= set(re.findall(r"\b(srrg2b_[a-z_0-9]+)\s*\(", open(os.path.join(ROOT, "adapters", name)).read()))
    used = {u for u in used if not u.startswith("srrg2b_adapters")}
    assert used and used <= declared, used - declared


@pytest.fixture(scope="module")
def detector_binary(tmp_path_factory):
    """adapters/multi_loop_detector_b200.h driven over the stub SLAM surface, linked against libsrrg2b.so."""
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build_cuda()
    exe = str(tmp_path_factory.mktemp("adapters") / "adapter_detector_main")
    libdir = os.path.dirname(G.LIB)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "adapters", "stubs"), "-I", os.path.join(ROOT, "adapters"),
                           os.path.join(ROOT, "tests", "cpp", "adapter_detector_main.cpp"), "-o", exe,
                           "-L", libdir, "-lsrrg2b", "-Wl,-rpath," + libdir])
    return exe


def test_loop_detector_adapter_links_and_fails_loudly_without_a_gpu(detector_binary):
    """On a machine without a CUDA device compute() throws instead of falling back to anything."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([detector_binary], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_loop_detector_adapter_detects_the_matching_map(detector_binary):
    """Source map = an L-shaped corner; candidates: the same corner 3 cm off (accepted: every point an inlier) and a
    corner 4 m away (the aligner drops it: no correspondences).  Reference flow:
    R/registration/loop_detector/multi_loop_detector_brute_force_impl.cpp:63-133."""
    r = subprocess.run([detector_binary], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "attempted 2 detected 1", lines
    assert lines[1].startswith("verdict 0 inliers 3000 correspondences 3000"), lines
    assert lines[2].startswith("verdict 1"), lines
    assert lines[-1] == "closure -> map 1", lines


@pytest.fixture(scope="module")
def tracker_binary(tmp_path_factory):
    """adapters/scene_b200.h + multi_aligner_b200.h: one tracker frame (clip -> align -> merge) over the stub surface."""
    import __graft_entry__ as G
    if not os.path.exists(G.LIB):
        G.build_cuda()
    exe = str(tmp_path_factory.mktemp("adapters") / "adapter_tracker_main")
    libdir = os.path.dirname(G.LIB)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "adapters", "stubs"), "-I", os.path.join(ROOT, "adapters"),
                           os.path.join(ROOT, "tests", "cpp", "adapter_tracker_main.cpp"), "-o", exe,
                           "-L", libdir, "-lsrrg2b", "-Wl,-rpath," + libdir])
    return exe


def test_scene_adapters_link_and_fail_loudly_without_a_gpu(tracker_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([tracker_binary], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_scene_adapters_run_a_tracker_frame(tracker_binary):
    """Local map = two 12 m walls (6000 points), scan = their first 8 m, 2.5 cm off.  The range clipper (9 m) keeps the
    4502 map points within range, the aligner -- its moving cloud resident, only the scan uploaded -- recovers the offset
    with every correspondence an inlier, the merger merges 1999 scan points and appends the remaining one
    (R/trackers/multi_tracker_impl.cpp:82-138, R/trackers/tracker_slice_processor_impl.cpp:159-205)."""
    r = subprocess.run([tracker_binary], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "clipped 4502 status 1", lines
    assert lines[1] == "aligner status 0 correspondences 4247 inliers 4247", lines
    assert lines[2] == "merged 1999 added 1 scene 6000 -> 6001", lines
    assert lines[3] == "map in scan: tx 0.0200 ty -0.0150", lines
