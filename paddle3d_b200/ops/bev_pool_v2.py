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
