"""the dense part of CenterPoint between the BEV tensor and the postprocess —
SecondBackbone (backbones/second_backbone.py:72-120), SecondFPN (necks/second_fpn.py:99-160, use_conv_for_no_stride)
and CenterHead (detection/centerpoint/center_head.py:43-220) — as a chain of `ops.dense_conv.dense_conv2d` launches
with BatchNorm folded into the conv epilogue.  Same constructor vocabulary as the reference's yml
(configs/centerpoint/centerpoint_voxels_0075voxel_nuscenes_10sweep.yml:127-162).  Parity-green against the CPU
reference (tests/test_gpu_dense.py); not part of the default bench frame yet."""
import numpy as np
import torch

from .ops import dense_conv as dc

COMMON_HEADS = (("reg", 2), ("height", 1), ("dim", 3), ("rot", 2), ("vel", 2))  # yml:157-162, then hm per task


class _Conv:
    """Conv2D / Conv2DTranspose (+ BatchNorm2D eval) (+ ReLU) with seeded parameters."""

    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False, bn_eps=None, relu=True, up=1):
        self.cin, self.cout, self.k, self.stride, self.padding, self.up = cin, cout, k, stride, padding, up
        self.has_bias, self.bn_eps, self.relu = bias, bn_eps, relu
        self.n_tile = dc.n_tile_for(cout)
        self.np = None
        self.dev = None

    def init(self, rng, device=None, randomize_bn=False, bias_value=None):
        """Seeded parameters (numpy); with a device also the packed tensor-core image and the folded epilogue."""
        cin, cout, k = self.cin, self.cout, self.k
        bound = 1.0 / np.sqrt(cin * k * k)  # build_conv_layer "uniform" (second_backbone.py:43-48)
        shape = (cin, cout, k, k) if self.up > 1 else (cout, cin, k, k)
        w = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        b = None
        if self.has_bias:
            b = (np.full(cout, bias_value, np.float32) if bias_value is not None
                 else rng.uniform(-bound, bound, size=cout).astype(np.float32))
        p = dict(weight=w, bias=b, stride=self.stride, padding=self.padding, up=self.up, relu=self.relu, bn=None)
        if self.bn_eps is not None:
            if randomize_bn:
                g, bt = rng.uniform(0.5, 1.5, cout), rng.uniform(-0.2, 0.2, cout)
                m, v = rng.uniform(-0.1, 0.1, cout), rng.uniform(0.5, 1.5, cout)
            else:
                g, bt, m, v = np.ones(cout), np.zeros(cout), np.zeros(cout), np.ones(cout)
            p["bn"] = dict(gamma=g.astype(np.float32), beta=bt.astype(np.float32), mean=m.astype(np.float32),
                           var=v.astype(np.float32), eps=self.bn_eps)
        self.np = p
        # fold: y = conv * s + ((bias - mean) * s + beta), s = gamma / sqrt(var + eps)   (fp64 on the host)
        s = np.ones(cout)
        t = np.zeros(cout) if b is None else b.astype(np.float64)
        if p["bn"] is not None:
            bn = p["bn"]
            s = bn["gamma"].astype(np.float64) / np.sqrt(bn["var"].astype(np.float64) + bn["eps"])
            t = (t - bn["mean"]) * s + bn["beta"]
        if device is None:
            return self
        self.dev = dict(
            packed=(dc.pack_deconv_weight if self.up > 1 else dc.pack_conv_weight)(torch.from_numpy(w).to(device), self.n_tile),
            scale=torch.from_numpy(s.astype(np.float32)).to(device) if p["bn"] is not None else None,
            shift=torch.from_numpy(t.astype(np.float32)).to(device) if (p["bn"] is not None or b is not None) else None)
        return self

    def __call__(self, x_split, shape, **kw):
        d = self.dev
        return dc.dense_conv2d(x_split, shape, d["packed"], self.cout, self.n_tile, self.k, self.stride, self.padding, self.up,
                               d["scale"], d["shift"], self.relu, **kw)


class DenseRPNHead:
    def __init__(self, in_channels=256, out_channels=(128, 256), layer_nums=(5, 5), downsample_strides=(1, 2),
                 fpn_out_channels=(256, 256), upsample_strides=(1, 2), tasks=(1, 2, 2, 1, 2, 2), share_conv_channel=64):
        self.tasks = list(tasks)
        bn3, bn5 = 1e-3, 1e-5
        self.blocks = []
        cin = in_channels
        for cout, n, s in zip(out_channels, layer_nums, downsample_strides):
            blk = [_Conv(cin, cout, 3, s, 1, bn_eps=bn3)] + [_Conv(cout, cout, 3, 1, 1, bn_eps=bn3) for _ in range(n)]
            self.blocks.append(blk)
            cin = cout
        self.deblocks = []
        for ci, co, u in zip(out_channels, fpn_out_channels, upsample_strides):
            # use_conv_for_no_stride: stride 1 -> Conv2D k = 1; stride > 1 -> Conv2DTranspose k = s (second_fpn.py:118-139)
            self.deblocks.append(_Conv(ci, co, 1, 1, 0, bn_eps=bn3) if u == 1 else _Conv(ci, co, u, u, 0, bn_eps=bn3, up=u))
        self.fpn_channels = int(sum(fpn_out_channels))
        self.shared = _Conv(self.fpn_channels, share_conv_channel, 3, 1, 1, bias=True, bn_eps=bn5)
        self.heads = []  # per task: list of (name, ConvModule 64->64, final conv 64->classes)
        for ncls in self.tasks:
            hs = []
            for name, c in list(COMMON_HEADS) + [("hm", ncls)]:
                hs.append((name, _Conv(share_conv_channel, share_conv_channel, 3, 1, 1, bias=True, bn_eps=bn5),
                           _Conv(share_conv_channel, c, 3, 1, 1, bias=True, relu=False)))
            self.heads.append(hs)

    def all_convs(self):
        out = [c for blk in self.blocks for c in blk] + list(self.deblocks) + [self.shared]
        for hs in self.heads:
            for _, a, b in hs:
                out += [a, b]
        return out

    def init_weight(self, seed=0, device="cuda", randomize_bn=False):
        """device=None: numpy parameters only (enough for export_numpy / the CPU arm)."""
        rng = np.random.default_rng(seed)
        hm_finals = {id(b) for hs in self.heads for name, _, b in hs if name == "hm"}
        for c in self.all_convs():  # hm bias = -2.19 (center_head.py:113-117)
            c.init(rng, device, randomize_bn, bias_value=-2.19 if id(c) in hm_finals else None)
        return self

    def export_numpy(self):
        return dict(blocks=[[c.np for c in blk] for blk in self.blocks], deblocks=[c.np for c in self.deblocks],
                    shared=self.shared.np, heads=[[(n, a.np, b.np) for n, a, b in hs] for hs in self.heads])

    def forward(self, bev):
        """bev [B, C, H, W] fp32 -> dict name -> list (per task) of [B, k, H, W] fp32 tensors."""
        b, c, h, w = bev.shape
        x, shape = dc.nchw_to_pixel_split(bev), (b, h, w, c)
        feats = []
        for blk in self.blocks:
            for conv in blk:
                x, _, (b_, oh, ow) = conv(x, shape)
                shape = (b_, oh, ow, conv.cout)
            feats.append((x, shape))
        cat, c0, out_hw = None, 0, None
        for (f, fshape), de in zip(feats, self.deblocks):
            if cat is None:
                up = de.up
                out_hw = (fshape[1] * up, fshape[2] * up)
                cat = torch.empty((fshape[0] * out_hw[0] * out_hw[1], 2 * self.fpn_channels), dtype=torch.float32, device=bev.device)
            de(f, fshape, out_split=cat, out_channels=self.fpn_channels, out_c0=c0)
            c0 += de.cout
        shape = (b, out_hw[0], out_hw[1], self.fpn_channels)
        s, _, _ = self.shared(cat, shape)
        shape = (b, out_hw[0], out_hw[1], self.shared.cout)
        out = {}
        for hs in self.heads:
            for name, a, fin in hs:
                t, _, _ = a(s, shape)
                _, planes, _ = fin(t, shape, want_nchw=True)
                out.setdefault(name, []).append(planes)
        return out

    __call__ = forward

    # ---- CenterHead.predict_by_custom_op (center_head.py:294-339): marshal the per-task head tensors
    def predict_by_custom_op(self, preds, test_cfg, with_velocity=True, example=None, postprocess_fn=None):
        """preds: forward()'s dict name -> [tensor per task], or the reference's list of per-task dicts.  test_cfg: dict
        with the yml's keys (voxel_size, point_cloud_range, post_center_limit_range, down_ratio, score_threshold,
        nms_iou_threshold, nms_pre_max_size, nms_post_max_size).  Returns the reference's one-element list of
        {meta, box3d_lidar, label_preds, scores}."""
        if postprocess_fn is None:
            from .ops.centerpoint_postprocess import centerpoint_postprocess as postprocess_fn
        if isinstance(preds, dict):
            preds = [{k: v[t] for k, v in preds.items()} for t in range(len(self.tasks))]
        hm, reg, height, dim, vel, rot, num_classes, flag = [], [], [], [], [], [], [], 0
        for task_id, pd in enumerate(preds):
            for nc in self.tasks:  # as in the reference the list grows to T*T entries; the op reads the first T
                num_classes.append(flag)
                flag += nc
            hm.append(pd["hm"])
            reg.append(pd["reg"])
            height.append(pd["height"])
            dim.append(pd["dim"])
            vel.append(pd["vel"] if with_velocity else pd["reg"])
            rot.append(pd["rot"])
        bboxes, scores, labels = postprocess_fn(
            hm, reg, height, dim, vel, rot, test_cfg["voxel_size"], test_cfg["point_cloud_range"],
            test_cfg["post_center_limit_range"], num_classes, test_cfg["down_ratio"], test_cfg["score_threshold"],
            test_cfg["nms_iou_threshold"], test_cfg["nms_pre_max_size"], test_cfg["nms_post_max_size"], with_velocity)
        meta = None if not example or not example.get("meta") else example["meta"][0]
        return [{"meta": meta, "box3d_lidar": bboxes, "label_preds": labels, "scores": scores}]

    # ---- EXPERIMENTAL (never run on a GPU yet): the 36 + 36 head convs as two launches
    def _batched_params(self, device):
        """One Conv(64 -> 36 * 64) with the 36 ConvModules' weights / folded BN concatenated along Cout, and the final
        convs' weights as [groups][9][Cin][4] for p3d_head_final_conv."""
        if getattr(self, "_batched", None) is not None:
            return self._batched
        firsts = [a for hs in self.heads for _, a, _ in hs]
        finals = [(n, f) for hs in self.heads for n, _, f in hs]
        cin = firsts[0].cin
        big = _Conv(cin, sum(a.cout for a in firsts), 3, 1, 1, bias=True, bn_eps=firsts[0].bn_eps)
        w = np.concatenate([a.np["weight"] for a in firsts], 0)
        scale, shift = [], []
        for a in firsts:  # same folding as _Conv.init
            bn = a.np["bn"]
            s_ = bn["gamma"].astype(np.float64) / np.sqrt(bn["var"].astype(np.float64) + bn["eps"])
            scale.append(s_)
            shift.append((a.np["bias"].astype(np.float64) - bn["mean"]) * s_ + bn["beta"])
        big.dev = dict(packed=dc.pack_conv_weight(torch.from_numpy(w).to(device), big.n_tile),
                       scale=torch.from_numpy(np.concatenate(scale).astype(np.float32)).to(device),
                       shift=torch.from_numpy(np.concatenate(shift).astype(np.float32)).to(device))
        groups = len(finals)
        fw = np.zeros((groups, 9, cin, 4), np.float32)
        fb = np.zeros((groups, 4), np.float32)
        plane0, cnt, p0 = [], [], 0
        for g, (_, f) in enumerate(finals):
            k = f.cout
            fw[g, :, :, :k] = f.np["weight"].transpose(2, 3, 1, 0).reshape(9, cin, k)  # [k, cin, 3, 3] -> [tap][cin][k]
            fb[g, :k] = f.np["bias"]
            plane0.append(p0)
            cnt.append(k)
            p0 += k
        self._batched = dict(big=big, fw=torch.from_numpy(fw).to(device), fb=torch.from_numpy(fb).to(device),
                             plane0=np.asarray(plane0, np.int32), cnt=np.asarray(cnt, np.int32), planes=p0,
                             names=[n for n, _ in finals])
        return self._batched

    def forward_batched(self, bev):
        """Same result as forward(); the 36 ConvModules run as one 64 -> 2304 conv and the 36 output convs as one
        grouped CUDA-core launch (p3d_head_final_conv)."""
        from ._lib import check, lib
        from ._mem import ptr, stream
        b, c, h, w = bev.shape
        x, shape = dc.nchw_to_pixel_split(bev), (b, h, w, c)
        feats = []
        for blk in self.blocks:
            for conv in blk:
                x, _, (b_, oh, ow) = conv(x, shape)
                shape = (b_, oh, ow, conv.cout)
            feats.append((x, shape))
        cat, c0, out_hw = None, 0, None
        for (f, fshape), de in zip(feats, self.deblocks):
            if cat is None:
                out_hw = (fshape[1] * de.up, fshape[2] * de.up)
                cat = torch.empty((fshape[0] * out_hw[0] * out_hw[1], 2 * self.fpn_channels), dtype=torch.float32, device=bev.device)
            de(f, fshape, out_split=cat, out_channels=self.fpn_channels, out_c0=c0)
            c0 += de.cout
        H, W = out_hw
        s, _, _ = self.shared(cat, (b, H, W, self.fpn_channels))
        bp = self._batched_params(bev.device)
        big = bp["big"]
        mid, _, _ = big(s, (b, H, W, self.shared.cout))  # [B*H*W][2][36 * 64]
        planes = torch.empty((b, bp["planes"], H, W), dtype=torch.float32, device=bev.device)
        check(lib().p3d_head_final_conv(ptr(mid), b, H, W, big.cout, self.shared.cout, len(bp["cnt"]), ptr(bp["fw"]),
                                        ptr(bp["fb"]), bp["plane0"].ctypes.data, bp["cnt"].ctypes.data, bp["planes"],
                                        ptr(planes), stream(bev.device)), "head_final_conv")
        out = {}
        for name, p0, k in zip(bp["names"], bp["plane0"], bp["cnt"]):
            out.setdefault(name, []).append(planes[:, int(p0):int(p0) + int(k)])
        return out
