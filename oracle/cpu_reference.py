"""CPU arm of the benchmark (bench.py `cpu_baseline` and `--impl reference`): one CenterPoint hot-path
frame on the host cores, through the test oracle (oracle/p3d_oracle.c, OpenMP) and — for the voxelizer —
the reference's own CPU op when oracle/_ref was built.  Nothing in the CUDA product path imports this."""
import time

import numpy as np


class CpuFrame:
    def __init__(self, cfg, weights, head, test_cfg, label_offsets, use_ref_voxelizer=True, dense_weights=None):
        """head: resident head tensors (dict name -> list per task) fed to the postprocess, or - when dense_weights
        (DenseRPNHead.export_numpy()) is given - ignored: the dense RPN / neck / CenterHead runs on the BEV tensor."""
        import oracle
        self.o = oracle
        self.cfg, self.w, self.head, self.tc, self.off = cfg, weights, head, test_cfg, label_offsets
        self.dense = CpuDenseHead(dense_weights) if dense_weights is not None else None
        self.use_ref = use_ref_voxelizer and oracle.ref_lib("cpu") is not None
        g = oracle.grid_size(cfg["voxel_size"], cfg["point_cloud_range"])
        self.sparse_shape = [g[2] + 1, g[1], g[0]]
        self.pairs = None

    def _conv(self, l, c, f, sp, subm):
        oc, of, osp, pairs = self.o.sparse_conv3d(c, f, 1, sp, l["weight"], l["stride"], l["padding"], subm)
        if l.get("bias") is not None:
            of = of + l["bias"]
        self.pairs.append(pairs)
        return oc, of, osp

    def _bn(self, l, f, relu=True, residual=None):
        return self.o.bn_relu(f, l["gamma"], l["beta"], l["mean"], l["var"], l["eps"], relu=relu, residual=residual)

    def _block(self, b, c, f, sp):
        _, o, _ = self._conv(b["conv1"], c, f, sp, True)
        o = self._bn(b["bn1"], o)
        _, o, _ = self._conv(b["conv2"], c, o, sp, True)
        return self._bn(b["bn2"], o, residual=f)

    def run(self, points):
        cfg, o = self.cfg, self.o
        t = {}
        t0 = time.perf_counter()
        vox = o.ref_hard_voxelize_cpu if self.use_ref else o.hard_voxelize
        v, c, n, nv = vox(points, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_points"], cfg["max_voxels"])
        k = int(nv[0])
        t["voxelize"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        f = o.voxel_mean(v, n, k)
        coors = np.concatenate([np.zeros((k, 1), np.int32), c[:k]], 1)
        t["voxel_mean"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.pairs = []
        w = self.w
        sp = self.sparse_shape
        c, f, _ = self._conv(w["conv_input"]["conv"], coors, f, sp, True)
        f = self._bn(w["conv_input"]["bn"], f)
        for b in w["blocks0"]:
            f = self._block(b, c, f, sp)
        for st in w["stages"]:
            c, f, sp = self._conv(st["down"]["conv"], c, f, sp, False)
            f = self._bn(st["down"]["bn"], f)
            for b in st["blocks"]:
                f = self._block(b, c, f, sp)
        c, f, sp = self._conv(w["extra"]["conv"], c, f, sp, False)
        f = self._bn(w["extra"]["bn"], f)
        t["sparse_backbone"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        bev = o.sparse_to_dense_bev(c, f, 1, sp)
        t["to_dense"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        h = self.head
        if self.dense is not None:
            h = self.dense.run(bev)
            t["dense_head"] = time.perf_counter() - t0
            t0 = time.perf_counter()
        tc = self.tc
        boxes, scores, labels, _ = o.centerpoint_postprocess(
            h["hm"], h["reg"], h["height"], h["dim"], h["vel"], h["rot"], cfg["voxel_size"][:2], cfg["point_cloud_range"],
            tc["post_center_limit_range"], self.off, tc["down_ratio"], tc["score_threshold"], tc["nms_iou_threshold"],
            tc["nms_pre_max_size"], tc["nms_post_max_size"], True)
        t["postprocess"] = time.perf_counter() - t0
        return dict(bev=bev, boxes=boxes, scores=scores, labels=labels, times=t, num_voxels=k, pairs=list(self.pairs),
                    head=h if self.dense is not None else None)


class CpuDenseHead:
    """CPU arm of the dense RPN / neck / head (dense_head.DenseRPNHead.export_numpy weights) through the oracle."""

    def __init__(self, weights):
        import oracle
        self.o, self.w = oracle, weights

    def _conv(self, l, x):
        o = self.o
        if l["up"] > 1:
            y = o.deconv2d(x, l["weight"], l["bias"], l["up"])
        else:
            y = o.conv2d(x, l["weight"], l["bias"], l["stride"], l["padding"])
        if l["bn"] is not None:
            bn = l["bn"]
            return o.bn2d_relu(y, bn["gamma"], bn["beta"], bn["mean"], bn["var"], bn["eps"], relu=l["relu"])
        return np.maximum(y, 0.0) if l["relu"] else y

    def run(self, bev):
        w, x, feats = self.w, bev, []
        for blk in w["blocks"]:
            for l in blk:
                x = self._conv(l, x)
            feats.append(x)
        cat = np.concatenate([self._conv(l, f) for l, f in zip(w["deblocks"], feats)], axis=1)
        s = self._conv(w["shared"], cat)
        out = {}
        for hs in w["heads"]:
            for name, a, fin in hs:
                out.setdefault(name, []).append(self._conv(fin, self._conv(a, s)))
        return out
