"""No-GPU checks of the boundary: the shared library loads, exports every symbol include/srrg2b.h
declares, refuses to compute without a CUDA device (no CPU fallback), and the ctypes mirrors have
the layout of the C structs."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(capi):
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build_cuda()
    return C.CDLL(capi.LIB_PATH)


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "srrg2b.h")).read()
    return sorted(set(re.findall(r"\b(srrg2b_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib, capi):
    declared = _declared_symbols()
    assert len(declared) >= 18
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert sorted(capi.EXPORTED_SYMBOLS) == declared


def test_version_and_no_cpu_fallback(lib, capi):
    assert lib.srrg2b_version() == 100
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path cannot be exercised")
    h = C.c_void_p()
    assert lib.srrg2b_ctx_create(3, 0, C.byref(h)) == capi.ERR_CUDA
    assert not h.value
    with pytest.raises(capi.Srrg2bError):
        capi.Context(3)


def test_struct_layouts_match_the_header(capi, tmp_path):
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "srrg2b.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(srrg2b_cloud),sizeof(srrg2b_finder_params),sizeof(srrg2b_factor_params),'
                   'sizeof(srrg2b_slice),sizeof(srrg2b_iter_stats),sizeof(srrg2b_aligner_params));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirrors = [capi.Cloud, capi.FinderParams, capi.FactorParams, capi.Slice, capi.IterStats, capi.AlignerParams]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_oracle_structs_match_its_header(oracle, tmp_path):
    src = tmp_path / "osizes.c"
    src.write_text('#include <stdio.h>\n#include "srrg2b_oracle.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(orc_finder_params),sizeof(orc_factor_params),sizeof(orc_slice),sizeof(orc_iter_stats),'
                   'sizeof(orc_aligner_params));return 0;}\n')
    exe = tmp_path / "osizes"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "oracle"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirrors = [oracle.FinderParams, oracle.FactorParams, oracle.Slice, oracle.IterStats, oracle.AlignerParams]
    assert sizes == [C.sizeof(m) for m in mirrors]
