"""`paddle3d.ops.bev_pool_v2` mirror — op `bev_pool_v2` (paddle3d/ops/bev_pool_v2/bev_pool.cc:111-118)."""
import torch

from .._lib import check, lib
from .._mem import ptr, require_cuda, stream


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts, bev_feat_shape):
    """Argument order as the reference op (lengths BEFORE starts; call site bevdet_transformer.py:44-46).
    depth [B*N,D,H,W] fp32, feat [B*N,H,W,C] fp32, ranks int32, bev_feat_shape (B, Y, X, C) -> out fp32."""
    depth = require_cuda(depth, "depth", torch.float32)
    feat = require_cuda(feat, "feat", torch.float32)
    rd = require_cuda(ranks_depth, "ranks_depth", torch.int32)
    rf = require_cuda(ranks_feat, "ranks_feat", torch.int32)
    rb = require_cuda(ranks_bev, "ranks_bev", torch.int32)
    il = require_cuda(interval_lengths, "interval_lengths", torch.int32)
    is_ = require_cuda(interval_starts, "interval_starts", torch.int32)
    c = feat.shape[3]  # bev_pool.cc:36
    out = torch.empty(tuple(int(s) for s in bev_feat_shape), dtype=torch.float32, device=feat.device)
    check(lib().p3d_bev_pool_v2(ptr(depth), ptr(feat), ptr(rd), ptr(rf), ptr(rb), ptr(il), ptr(is_), il.shape[0], c,
                                ptr(out), out.numel(), stream(feat.device)), "bev_pool_v2")
    return out


def voxel_pooling_prepare_v2(coor, grid_lower_bound, grid_interval, grid_size):
    """LSSViewTransformer.voxel_pooling_prepare_v2 (bevdet_transformer.py:230-274) on the device, no host sync:
    coor [B, N, D, H, W, 3] fp32 -> (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths, counts);
    the five rank arrays are int32 tensors of capacity B*N*D*H*W whose first counts[0] (ranks) / counts[1] (intervals)
    entries are valid (`trim` slices them like the reference returns them, at the cost of one D2H read)."""
    from .._lib import host_floats, host_ints
    from .._mem import workspace
    coor = require_cuda(coor, "coor", torch.float32)
    B, N, D, H, W, three = coor.shape
    assert three == 3
    n = B * N * D * H * W
    dev = coor.device
    outs = [torch.empty((n,), dtype=torch.int32, device=dev) for _ in range(5)]
    counts = torch.empty((2,), dtype=torch.int32, device=dev)
    L = lib()
    wsb = L.p3d_bev_pool_prepare_workspace_bytes(n)
    ws = workspace(wsb, dev, "bev_pool_prepare")
    check(L.p3d_bev_pool_prepare(ptr(coor), B, N, D, H, W, host_floats(grid_lower_bound), host_floats(grid_interval),
                                 host_ints(grid_size), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), ptr(outs[4]),
                                 ptr(counts), ptr(ws), wsb, stream(dev)), "bev_pool_prepare")
    return outs[0], outs[1], outs[2], outs[3], outs[4], counts


def trim(prepared):
    """Slice the capacity-sized outputs of voxel_pooling_prepare_v2 to their valid lengths (one D2H read), in the
    reference's return order (ranks_bev, ranks_depth, ranks_feat, interval_starts, interval_lengths)."""
    rb, rd, rf, st, ln, counts = prepared
    k, m = [int(v) for v in counts.cpu()]
    if k == 0 or m == 0:
        return None, None, None, None, None
    return rb[:k], rd[:k], rf[:k], st[:m], ln[:m]
