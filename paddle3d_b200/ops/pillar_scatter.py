"""Dense canvas writer: PointPillarsScatter (middle_encoders/pillar_scatter.py:57-105) and
SparseResNet3D's to_dense + transpose + reshape (sparse_resnet.py:202-206)."""
import torch

from .._lib import check, lib
from .._mem import ptr, require_cuda, stream, workspace


def _scatter(feats, coords, batch_size, D, ny, nx, use_z, num=None):
    feats = require_cuda(feats, "voxel_features", torch.float32)
    coords = require_cuda(coords, "coords", torch.int32)
    n, C = feats.shape
    dev = feats.device
    out = torch.empty((batch_size, C, D, ny, nx), dtype=torch.float32, device=dev)
    L = lib()
    ws = workspace(L.p3d_scatter_dense_workspace_bytes(batch_size, D, ny, nx), dev, "scatter")
    nump = ptr(require_cuda(num, "num", torch.int32)) if num is not None else ptr(None)
    check(L.p3d_scatter_dense(ptr(feats), ptr(coords), nump, n, C, batch_size, D, ny, nx, int(use_z), ptr(out),
                              ptr(ws), ws.numel(), stream(dev)), "scatter_dense")
    return out


def pillar_scatter(voxel_features, coords, batch_size, ny, nx, num=None):
    """[n, C] features + [n, 4] (b, z, y, x) int32 coords -> [batch, C, ny, nx] canvas."""
    return _scatter(voxel_features, coords, batch_size, 1, ny, nx, False, num).view(batch_size, -1, ny, nx)


def sparse_to_dense_bev(feats, coords, batch_size, spatial_shape, num=None):
    """[n, C] + [n, 4] -> [batch, C*D, H, W] (out.to_dense().transpose(0,4,1,2,3).reshape(N, C*D, H, W))."""
    D, H, W = [int(s) for s in spatial_shape]
    out = _scatter(feats, coords, batch_size, D, H, W, True, num)
    return out.view(batch_size, -1, H, W)
