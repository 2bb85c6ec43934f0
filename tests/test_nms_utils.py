"""CPU: the Python NMS callers (paddle3d_b200/ops/nms_utils.py) with the oracle's NMS / overlap injected
for the GPU ops, against a plain numpy restatement of the reference's index logic
(layer_libs.py:210-249, model_nms_utils.py:20-68, iou3d_nms_utils.py:25-57)."""
import numpy as np
import torch

from paddle3d_b200 import synth
from paddle3d_b200.ops import nms_utils


def _nms_fn(oracle_mod):
    def fn(boxes, thresh):
        keep, n = oracle_mod.nms(boxes.numpy(), thresh)
        return torch.from_numpy(np.asarray(keep, np.int64)), torch.tensor([n], dtype=torch.int64)
    return fn


def test_rotate_nms_pcdet(oracle_mod):
    rng = np.random.default_rng(0)
    boxes = synth.random_boxes(300, 5)          # (x, y, z, w, l, h, theta)
    scores = rng.uniform(size=300).astype(np.float32)
    scores[10] = scores[20]                     # a tie: lower index first
    got = nms_utils.rotate_nms_pcdet(torch.from_numpy(boxes), torch.from_numpy(scores), 0.2, 200, 50, nms_fn=_nms_fn(oracle_mod))
    b = boxes[:, [0, 1, 2, 4, 3, 5, 6]].copy()
    b[:, -1] = -b[:, -1] - np.float32(np.pi / 2)
    order = np.argsort(-scores, kind="stable")[:200]
    keep, n = oracle_mod.nms(b[order], 0.2)
    want = order[keep[:n]][:50]
    assert np.array_equal(got.numpy(), want)
    assert len(got) <= 50 and len(set(got.tolist())) == len(got)
    one = nms_utils.rotate_nms_pcdet(torch.from_numpy(boxes[:1]), torch.from_numpy(scores[:1]), 0.2, nms_fn=_nms_fn(oracle_mod))
    assert one.tolist() == [0]


def test_class_agnostic_nms(oracle_mod):
    rng = np.random.default_rng(1)
    boxes = synth.random_boxes(200, 6)
    scores = rng.uniform(size=200).astype(np.float32)
    labels = rng.integers(0, 3, size=200).astype(np.int64)
    cfg = dict(nms_pre_maxsize=150, nms_post_maxsize=40, nms_thresh=0.3)
    s, l, b = nms_utils.class_agnostic_nms(torch.from_numpy(scores), torch.from_numpy(boxes), torch.from_numpy(labels), cfg,
                                           score_thresh=0.25, nms_fn=_nms_fn(oracle_mod))
    idx = np.nonzero(scores >= 0.25)[0]
    order = idx[np.argsort(-scores[idx], kind="stable")][:150]
    keep, n = oracle_mod.nms(boxes[order], 0.3)
    sel = order[keep[:n]][:40]
    assert np.array_equal(s.numpy(), scores[sel]) and np.array_equal(l.numpy(), labels[sel])
    assert np.array_equal(b.numpy(), boxes[sel])
    s, l, b = nms_utils.class_agnostic_nms(torch.from_numpy(scores), torch.from_numpy(boxes), torch.from_numpy(labels), cfg,
                                           score_thresh=2.0, nms_fn=_nms_fn(oracle_mod))
    assert s.tolist() == [-1.0] and l.tolist() == [-1] and b.shape == (1, 7) and float(b.abs().sum()) == 0.0


def test_boxes_iou3d(oracle_mod):
    a, b = synth.random_boxes(40, 7), synth.random_boxes(30, 8)
    b[:5] = a[:5]
    fn = lambda x, y: torch.from_numpy(oracle_mod.boxes_overlap_bev(x.numpy(), y.numpy()))  # noqa: E731
    got = nms_utils.boxes_iou3d_gpu(torch.from_numpy(a), torch.from_numpy(b), overlap_fn=fn).numpy()
    assert got.shape == (40, 30) and (got >= 0).all() and (got <= 1 + 1e-5).all()
    np.testing.assert_allclose(np.diag(got[:5, :5]), 1.0, rtol=1e-4)   # identical boxes
    ov = oracle_mod.boxes_overlap_bev(a, b)
    h = np.clip(np.minimum(a[:, None, 2] + a[:, None, 5] / 2, b[None, :, 2] + b[None, :, 5] / 2)
                - np.maximum(a[:, None, 2] - a[:, None, 5] / 2, b[None, :, 2] - b[None, :, 5] / 2), 0, None)
    va, vb = (a[:, 3] * a[:, 4] * a[:, 5])[:, None], (b[:, 3] * b[:, 4] * b[:, 5])[None, :]
    np.testing.assert_allclose(got, ov * h / np.clip(va + vb - ov * h, 1e-6, None), rtol=1e-5, atol=1e-7)


def test_predict_by_custom_op_marshalling():
    """CenterHead.predict_by_custom_op (center_head.py:294-339): argument order, the num_classes prefix
    list (first T entries = label offsets), vel falling back to reg, the returned record."""
    from paddle3d_b200.dense_head import DenseRPNHead
    net = DenseRPNHead(in_channels=32, out_channels=(32,), layer_nums=(0,), downsample_strides=(1,), fpn_out_channels=(32,),
                       upsample_strides=(1,), tasks=(1, 2, 2), share_conv_channel=32, with_velocity=False)
    preds = {k: [("%s%d" % (k, t)) for t in range(3)] for k in ("hm", "reg", "height", "dim", "vel", "rot")}
    seen = {}

    def fake(*args):
        seen["args"] = args
        return "B", "S", "L"

    # the yml's nested form (test_cfg.nms.*), positional order of the reference call site (centerpoint.py:163)
    cfg = dict(voxel_size=[0.075, 0.075], point_cloud_range=[-54, -54, -5, 54, 54, 3], post_center_limit_range=[-61.2] * 3 + [61.2] * 3,
               down_ratio=8, score_threshold=0.1, nms=dict(nms_iou_threshold=0.2, nms_pre_max_size=1000, nms_post_max_size=83))
    out = net.predict_by_custom_op({"meta": ["m0"]}, preds, cfg, postprocess_fn=fake)
    a = seen["args"]
    assert a[0] == ["hm0", "hm1", "hm2"] and a[4] == ["reg0", "reg1", "reg2"] and a[5] == ["rot0", "rot1", "rot2"]
    assert a[9][:3] == synth.label_offsets((1, 2, 2)) == [0, 1, 3] and len(a[9]) == 9
    assert a[6] == cfg["voxel_size"] and a[10] == 8 and a[13:16] == (1000, 83, False)
    assert out == [{"meta": "m0", "box3d_lidar": "B", "label_preds": "L", "scores": "S"}]
