"""`paddle3d.ops.bev_pool_v2_backward` mirror — op `bev_pool_v2_bkwd`
(paddle3d/ops/bev_pool_v2_backward/bev_pool_bkwd.cc:75-80; call site bevdet_transformer.py:69-78)."""
import torch

from .._lib import check, lib
from .._mem import ptr, require_cuda, stream


def bev_pool_v2_bkwd(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts):
    out_grad = require_cuda(out_grad, "out_grad", torch.float32)
    depth = require_cuda(depth, "depth", torch.float32)
    feat = require_cuda(feat, "feat", torch.float32)
    rd = require_cuda(ranks_depth, "ranks_depth", torch.int32)
    rf = require_cuda(ranks_feat, "ranks_feat", torch.int32)
    rb = require_cuda(ranks_bev, "ranks_bev", torch.int32)
    il = require_cuda(interval_lengths, "interval_lengths", torch.int32)
    is_ = require_cuda(interval_starts, "interval_starts", torch.int32)
    c = out_grad.shape[3]
    dg = torch.empty_like(depth)
    fg = torch.empty_like(feat)
    check(lib().p3d_bev_pool_v2_bkwd(ptr(out_grad), ptr(depth), ptr(feat), ptr(rd), ptr(rf), ptr(rb), ptr(il),
                                     ptr(is_), il.shape[0], c, ptr(dg), dg.numel(), ptr(fg), fg.numel(),
                                     stream(feat.device)), "bev_pool_v2_bkwd")
    return dg, fg
