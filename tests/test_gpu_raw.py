"""The torch-free front end (paddle3d_b200/raw.py: ctypes + libcudart + numpy) on the GPU box, in a FRESH interpreter so
that `torch` demonstrably never gets imported: hard_voxelize bit-exact, rotated IoU and NMS keep list vs the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
import numpy as np
sys.path.insert(0, %r)
import paddle3d_b200.raw as raw
import oracle
from paddle3d_b200 import synth
assert "torch" not in sys.modules, "paddle3d_b200.raw must not import torch"
cfg = dict(synth.C3, point_cloud_range=[-20.0, -20.0, -5.0, 20.0, 20.0, 3.0])
pts = synth.lidar_cloud(cfg, 5, num_points=50000)
P, V = 10, 30000
got = raw.hard_voxelize(pts, cfg["voxel_size"], cfg["point_cloud_range"], P, V)
want = oracle.hard_voxelize(pts, cfg["voxel_size"], cfg["point_cloud_range"], P, V)
for g, w in zip(got, want):
    assert np.array_equal(g, w), "hard_voxelize mismatch"
boxes = synth.random_boxes(300, 2)
iou = raw.boxes_iou_bev(boxes[:40], boxes[40:100])
np.testing.assert_allclose(iou, oracle.boxes_iou_bev(boxes[:40], boxes[40:100]), rtol=1e-4, atol=1e-6)
keep, num = raw.nms_gpu(boxes, 0.2)
wk, wn = oracle.nms(boxes, 0.2)
assert num == wn and np.array_equal(keep[:wn], wk[:wn]), "nms keep list mismatch"
assert "torch" not in sys.modules
print("raw ok", int(want[3][0]), wn)
"""


def test_raw_front_end_without_torch():
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "raw ok" in r.stdout
