"""CPU suite: the C-ABI library loads and exports every symbol include/p3d_b200.h declares;
the host mirror refuses to run without CUDA tensors (no fallback).  No compute calls."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "p3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p3d_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from paddle3d_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
        assert n in _lib.SIGNATURES, "no ctypes signature for %s" % n
    assert lib.p3d_abi_version() == 1


def test_status_strings_and_size_queries():
    from paddle3d_b200 import _lib
    L = _lib.lib()
    assert L.p3d_status_string(0) == b"ok"
    assert b"workspace" in L.p3d_status_string(-2)
    assert L.p3d_hard_voxelize_workspace_bytes(300000, 10, 160000) > 8 * 2 ** 20
    assert L.p3d_hard_voxelize_workspace_bytes(-1, 10, 10) == 0
    assert L.p3d_nms_workspace_bytes(1000) >= 1000 * 16 * 8
    assert L.p3d_scatter_dense_workspace_bytes(1, 1, 496, 432) >= 496 * 432 * 4
    assert L.p3d_centerpoint_postprocess_workspace_bytes(6, 180, 180, 1000, 83) > 0
    assert L.p3d_sparse_rulebook_workspace_bytes(160000, 640000) > 0


def test_host_mirror_rejects_cpu_tensors():
    import torch
    from paddle3d_b200 import _lib
    from paddle3d_b200.ops import iou3d_nms, voxelize
    with pytest.raises(_lib.P3DError):
        voxelize.hard_voxelize(torch.zeros(10, 4), [0.16, 0.16, 4], [0, -39.68, -3, 69.12, 39.68, 1], 32, 100)
    with pytest.raises(_lib.P3DError):
        iou3d_nms.boxes_iou_bev_cpu(torch.zeros(2, 7), torch.zeros(2, 7))


def test_invalid_arguments_return_status_codes():
    from paddle3d_b200 import _lib
    L = _lib.lib()
    vs = _lib.host_floats([0.1, 0.1, 0.1])
    pcr = _lib.host_floats([0, 0, 0, 1, 1, 1])
    # null outputs -> invalid argument, before any CUDA call
    rc = L.p3d_hard_voxelize(None, 0, 4, vs, pcr, 4, 16, None, None, None, None, None, 0, None)
    assert rc == -1
    # grid too large for 31-bit cell ids -> unsupported
    big = _lib.host_floats([0, 0, 0, 1e5, 1e5, 1e3])
    tiny = _lib.host_floats([0.01, 0.01, 0.01])
    rc = L.p3d_hard_voxelize(None, 0, 4, tiny, big, 4, 16, None, None, None, None, None, 0, None)
    assert rc == -4
    assert L.p3d_nms(None, -1, 0.5, 0, None, None, None, 0, None) == -1


def test_paddle_glue_compiles():
    """paddle_ext/p3d_paddle_ops.cc (the PD_BUILD_OP registrations that bind the C ABI under PaddlePaddle) must at
    least compile against the stub extension header: PaddlePaddle itself is not installable here."""
    import subprocess
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "oracle", "stub"), "-I",
                        os.path.join(ROOT, "include"), os.path.join(ROOT, "paddle_ext", "p3d_paddle_ops.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "paddle_ext", "p3d_paddle_ops.cc")).read()
    for op in ("hard_voxelize", "boxes_iou_bev_gpu", "boxes_overlap_bev_gpu", "nms_gpu", "nms_normal_gpu",
               "centerpoint_postprocess", "bev_pool_v2", "bev_pool_v2_bkwd"):
        assert "PD_BUILD_OP(%s)" % op in src
