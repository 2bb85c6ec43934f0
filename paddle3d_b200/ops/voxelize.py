"""`paddle3d.ops.voxelize` mirror — op `hard_voxelize` (paddle3d/ops/voxel/voxelize_op.cc:183-191)."""
import torch

from .._lib import check, host_floats, lib
from .._mem import ptr, require_cuda, stream, workspace


def hard_voxelize(points, voxel_size, point_cloud_range, max_num_points_in_voxel, max_voxels):
    """Same signature and outputs as the reference op (call site voxelizers/voxelize.py:40-42):
    returns (voxels [V,P,F] fp32 zero padded, coords [V,3] int32 (z,y,x), num_points_per_voxel [V] int32,
    num_voxels [1] int32), all sized by max_voxels (HardInferShape, voxelize_op.cc:168-176)."""
    points = require_cuda(points, "points", torch.float32)
    if points.dim() != 2 or points.shape[1] < 3:
        raise ValueError("points must be [N, >=3]")
    n, f = points.shape
    dev = points.device
    P, V = int(max_num_points_in_voxel), int(max_voxels)
    voxels = torch.empty((V, P, f), dtype=torch.float32, device=dev)
    coords = torch.empty((V, 3), dtype=torch.int32, device=dev)
    npv = torch.empty((V,), dtype=torch.int32, device=dev)
    nv = torch.empty((1,), dtype=torch.int32, device=dev)
    L = lib()
    ws_bytes = L.p3d_hard_voxelize_workspace_bytes(n, P, V)
    ws = workspace(ws_bytes, dev, "voxelize")
    vs, pcr = host_floats(voxel_size), host_floats(point_cloud_range)
    check(L.p3d_hard_voxelize(ptr(points), n, f, vs, pcr, P, V, ptr(voxels), ptr(coords), ptr(npv), ptr(nv), ptr(ws),
                              ws.numel(), stream(dev)), "hard_voxelize")
    return voxels, coords, npv, nv


def voxelize_mean(points, voxel_size, point_cloud_range, max_num_points_in_voxel, max_voxels, batch_id=0):
    """Fused HardVoxelizer + VoxelMean front end: returns
    (mean [V,F], coors [V,4] (b,z,y,x), num_points_per_voxel [V], num_voxels [1]) without the padded tensor."""
    points = require_cuda(points, "points", torch.float32)
    n, f = points.shape
    dev = points.device
    P, V = int(max_num_points_in_voxel), int(max_voxels)
    mean = torch.empty((V, f), dtype=torch.float32, device=dev)
    coors = torch.empty((V, 4), dtype=torch.int32, device=dev)
    npv = torch.empty((V,), dtype=torch.int32, device=dev)
    nv = torch.empty((1,), dtype=torch.int32, device=dev)
    L = lib()
    ws = workspace(L.p3d_hard_voxelize_workspace_bytes(n, P, V), dev, "voxelize")
    check(L.p3d_voxelize_mean(ptr(points), n, f, host_floats(voxel_size), host_floats(point_cloud_range), P, V,
                              int(batch_id), ptr(mean), ptr(coors), ptr(npv), ptr(nv), ptr(ws), ws.numel(),
                              stream(dev)), "voxelize_mean")
    return mean, coors, npv, nv


def voxel_mean(voxels, num_points_per_voxel, num_voxels=None):
    """VoxelMean.forward (voxel_encoders/voxel_encoder.py:49-57) on the padded tensor."""
    voxels = require_cuda(voxels, "voxels", torch.float32)
    npv = require_cuda(num_points_per_voxel, "num_points_per_voxel", torch.int32)
    cap, P, F = voxels.shape
    out = torch.empty((cap, F), dtype=torch.float32, device=voxels.device)
    nvp = ptr(require_cuda(num_voxels, "num_voxels", torch.int32)) if num_voxels is not None else ptr(None)
    check(lib().p3d_voxel_mean(ptr(voxels), ptr(npv), nvp, cap, P, F, ptr(out), stream(voxels.device)), "voxel_mean")
    return out
