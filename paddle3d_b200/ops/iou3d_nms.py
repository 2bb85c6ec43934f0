"""`paddle3d.ops.iou3d_nms` mirror (registrations paddle3d/ops/iou3d_nms/iou3d_nms_api.cpp:73-108).
boxes are [n, 7] = x, y, z, dx, dy, dz, heading."""
import torch

from .._lib import check, lib
from .._mem import ptr, require_cuda, stream, workspace


def _pair(a, b, iou):
    a = require_cuda(a, "boxes_a", torch.float32)
    b = require_cuda(b, "boxes_b", torch.float32)
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    fn = lib().p3d_boxes_iou_bev if iou else lib().p3d_boxes_overlap_bev
    check(fn(ptr(a), a.shape[0], ptr(b), b.shape[0], ptr(out), stream(a.device)), "boxes_iou_bev" if iou else "boxes_overlap_bev")
    return out


def boxes_overlap_bev_gpu(boxes_a, boxes_b):
    """iou3d_nms.cpp:44-62"""
    return _pair(boxes_a, boxes_b, False)


def boxes_iou_bev_gpu(boxes_a, boxes_b):
    """iou3d_nms.cpp:64-84"""
    return _pair(boxes_a, boxes_b, True)


def boxes_iou_bev_cpu(boxes_a, boxes_b):
    """The reference registers a CPU kernel under this name (iou3d_cpu.cpp:241-264).  This build has no
    CPU path: GPU tensors are served by the CUDA kernel, CPU tensors are rejected loudly."""
    return _pair(boxes_a, boxes_b, True)


def _nms(boxes, thresh, normal, device_outputs):
    boxes = require_cuda(boxes, "boxes", torch.float32)
    n = boxes.shape[0]
    dev = boxes.device
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    L = lib()
    ws = workspace(L.p3d_nms_workspace_bytes(n), dev, "nms")
    check(L.p3d_nms(ptr(boxes), n, float(thresh), int(normal), ptr(keep), ptr(num), ptr(ws), ws.numel(), stream(dev)),
          "nms")
    keep = keep[:n]
    if device_outputs:
        return keep, num
    # the reference returns CPU tensors (iou3d_nms.cpp:89-92): one D2H copy here, none inside the op
    return keep.cpu(), num.cpu()


def nms_gpu(boxes, nms_overlap_thresh, device_outputs=False):
    """iou3d_nms.cpp:86-141: returns (keep int32 [N] with the first num_to_keep entries valid, num_to_keep [1])."""
    return _nms(boxes, nms_overlap_thresh, False, device_outputs)


def nms_normal_gpu(boxes, nms_overlap_thresh, device_outputs=False):
    """iou3d_nms.cpp:143-204 (axis-aligned IoU)."""
    return _nms(boxes, nms_overlap_thresh, True, device_outputs)
