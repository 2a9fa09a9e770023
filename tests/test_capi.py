"""CPU tests of the C-ABI boundary: the library builds for sm_100a, loads, exports every symbol include/p3p.h
declares, and its host-side argument checking / size queries work without a GPU (no compute call is made)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "p3p.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p3p_[a-z_0-9]+)\s*\(", src)))


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "p3p.h"\nint main(void){p3p_grid g; (void)g; return P3P_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o", str(tmp_path / "t.o")])


def test_library_exports_every_declared_symbol(built_lib):
    from pixelspointspolygons_b200 import _lib

    names = declared_functions()
    assert set(names) == set(_lib.EXPORTED), (names, _lib.EXPORTED)
    raw = C.CDLL(built_lib)
    for n in names:
        assert hasattr(raw, n), n
    assert _lib.lib().p3p_version() == 200


def test_library_targets_sm_100a_with_tcgen05(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass or "UTC" in sass  # tcgen05.mma
    assert "LDTM" in sass                                          # tcgen05.ld
    assert "UTMALDG" in sass                                       # TMA tensor loads (cp.async.bulk.tensor) of the convolution
    assert "FMNMX3" in sass                                        # 3-input max in the epilogue


def test_size_queries_and_argument_errors(built_lib):
    from pixelspointspolygons_b200 import _lib

    l = _lib.lib()
    g = _lib.make_grid((0, 0, 0), (224, 224, 100), (8, 8, 100), 64, 784, 28, 28)
    n = l.p3p_workspace_bytes(C.byref(g), 16, 1_600_000)
    # slots dominate: B * keys * M * 16 bytes
    assert n > 16 * 1597 * 64 * 16 and n < 64 * 1024 * 1024
    assert l.p3p_workspace_bytes(C.byref(g), 0, 0) >= 0
    assert l.p3p_pfn_blob_bytes(384) > 384 * 64 * 4
    # bad grid -> 0 bytes and a message
    bad = _lib.make_grid((0, 0, 0), (224, 224, 100), (0, 8, 100), 64, 784, 28, 28)
    assert l.p3p_workspace_bytes(C.byref(bad), 1, 10) == 0
    assert b"voxel_size" in l.p3p_last_error()
    # null pointers are rejected before any CUDA call
    rc = l.p3p_encode(None, 3, None, 1, 10, C.byref(g), None, 384, 1, None, 0, 0, 384, 0, 0, None, 0, None)
    assert rc == -1
    rc = l.p3p_voxelize(None, 2, None, 0, 0, C.byref(g), None, None, 0, None)
    assert rc == -1 and b"point_stride" in l.p3p_last_error()
    with pytest.raises(_lib.P3PError):
        _lib.check(rc, "p3p_voxelize")
    # the 3x3 convolution: blob size, operand type check, channel range check (before any CUDA call)
    assert l.p3p_conv3x3_blob_bytes(768, 384) == 384 * 9 * 768 * 2 + 384 * 4
    rc = l.p3p_conv3x3(None, 1, 28, 28, 768, None, 384, 3, 1, None, 1, 384, 0, None)
    assert rc == -1
    rc = l.p3p_patch_embed(None, 1, 3, 224, 224, 8, None, None, 384, 3, None, 0, 7, 384, 0, None)
    assert rc == -1 and b"layout" in l.p3p_last_error()
    huge = _lib.make_grid((0, 0, 0), (2240, 2240, 100), (8, 8, 100), 64, 784, 280, 280)
    assert l.p3p_workspace_bytes(C.byref(huge), 1, 10) == 0 and b"ranking budget" in l.p3p_last_error()
