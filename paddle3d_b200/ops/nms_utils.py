"""Python callers of the iou3d_nms ops, same names and argument meaning as the reference:
`rotate_nms_pcdet` (models/layers/layer_libs.py:210-249), `class_agnostic_nms` (models/common/model_nms_utils.py:20-68)
and `boxes_iou3d_gpu` (models/heads/roi_heads/target_assigner/iou3d_nms_utils.py:25-57).  Host-side index logic in
torch on whatever device the inputs live on; the box geometry runs in `ops.iou3d_nms` (GPU only).  Score ties sort by
ascending index (Paddle's argsort tie order is unspecified), as in `centerpoint_postprocess`.

`nms_fn` / `overlap_fn` exist for the CPU test-suite, which injects the oracle's NMS; product code leaves them None."""
import numpy as np
import torch


def _nms_gpu(boxes, thresh):
    from . import iou3d_nms
    return iou3d_nms.nms_gpu(boxes, thresh)


def _take_keep(keep, num_out, device):
    # the reference op returns `keep` / `num_out` as CPU tensors (iou3d_nms.cpp:89-92); index tensors must live with
    # the tensor they index
    n = int(num_out[0]) if torch.is_tensor(num_out) else int(num_out)
    return keep[:n].to(device=device, dtype=torch.long)


def rotate_nms_pcdet(boxes, scores, thresh, pre_max_size=None, post_max_size=None, nms_fn=None):
    """boxes [N, >=7] (x, y, z, w, l, h, ..., theta), scores [N] -> indices of the kept boxes, best first."""
    nms_fn = nms_fn or _nms_gpu
    cols = torch.tensor([0, 1, 2, 4, 3, 5, int(boxes.shape[-1]) - 1], device=boxes.device)
    b = boxes.index_select(-1, cols).clone()  # back to pcdet's (x, y, z, l, w, h, theta) convention
    b[:, -1] = -b[:, -1] - np.pi / 2
    order = torch.argsort(scores, dim=0, descending=True, stable=True)
    if pre_max_size is not None:
        order = order[:pre_max_size]
    b = b[order].reshape(-1, 7)
    keep, num_out = nms_fn(b.contiguous(), thresh)
    selected = order[_take_keep(keep, num_out, order.device)]
    if post_max_size is not None:
        selected = selected[:post_max_size]
    return selected


def class_agnostic_nms(box_scores, box_preds, label_preds, nms_config, score_thresh=None, nms_fn=None):
    """-> (selected_score, selected_label, selected_box); one fake row (-1, -1, zeros) when nothing passes the score
    threshold (model_nms_utils.py:45-53)."""
    nms_fn = nms_fn or _nms_gpu

    def nms(scores, boxes, labels):
        order = torch.argsort(scores, dim=0, descending=True, stable=True)[:nms_config["nms_pre_maxsize"]]
        boxes, scores, labels = boxes[order], scores[order], labels[order]
        keep, num_out = nms_fn(boxes.contiguous(), nms_config["nms_thresh"])
        sel = _take_keep(keep, num_out, boxes.device)[:nms_config["nms_post_maxsize"]]
        return scores[sel], labels[sel], boxes[sel]

    if score_thresh is None:
        return nms(box_scores, box_preds, label_preds)
    mask = box_scores >= score_thresh
    if not bool(mask.any()):
        dev = box_scores.device
        return (torch.tensor([-1.0], dtype=box_scores.dtype, device=dev),
                torch.tensor([-1.0], device=dev).to(label_preds.dtype),
                torch.zeros((1, 7), dtype=box_preds.dtype, device=dev))
    idx = torch.nonzero(mask).reshape(-1)
    return nms(box_scores[idx], box_preds[idx], label_preds[idx])


def boxes_iou3d_gpu(boxes_a, boxes_b, overlap_fn=None):
    """[N, 7] x [M, 7] (x, y, z, dx, dy, dz, heading) -> 3-D IoU [N, M]: BEV overlap x height overlap."""
    if overlap_fn is None:
        from . import iou3d_nms
        overlap_fn = iou3d_nms.boxes_overlap_bev_gpu
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_max = (boxes_a[:, 2] + boxes_a[:, 5] / 2).reshape(-1, 1)
    a_min = (boxes_a[:, 2] - boxes_a[:, 5] / 2).reshape(-1, 1)
    b_max = (boxes_b[:, 2] + boxes_b[:, 5] / 2).reshape(1, -1)
    b_min = (boxes_b[:, 2] - boxes_b[:, 5] / 2).reshape(1, -1)
    overlaps_bev = overlap_fn(boxes_a, boxes_b)
    overlaps_h = torch.clamp(torch.minimum(a_max, b_max) - torch.maximum(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).reshape(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).reshape(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)
