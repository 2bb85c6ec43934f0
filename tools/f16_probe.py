#!/usr/bin/env python
"""Timeline probe of the fp16-pair sparse-conv kernel (csrc/sparse_conv_f16.cu) on the real C3 frame.

Runs the SparseResNet3D backbone once eagerly with precision F16X3; every tensor-core layer launch writes the clock64
timeline of one CTA (p3d_debug_f16_timeline) into its own buffer.  Prints one JSON line per layer with the medians that
say which role is the bottleneck.    python tools/f16_probe.py [--cta 0] [--frame 0] > gpurun_out/f16_probe.jsonl
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from paddle3d_b200 import synth  # noqa: E402
from paddle3d_b200._lib import lib  # noqa: E402
from paddle3d_b200.ops import sparse_nn as sp  # noqa: E402
from paddle3d_b200.ops import voxelize as vox  # noqa: E402
from paddle3d_b200.pipeline import CenterPointHotPath  # noqa: E402


def med(a):
    a = [x for x in a if x is not None]
    return int(np.median(a)) if a else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cta", type=int, default=0)
    ap.add_argument("--frame", type=int, default=0)
    ap.add_argument("--geometry", default="0075")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = synth.C3 if args.geometry == "0075" else synth.C3_01
    pipe = CenterPointHotPath(cfg, dev, precision=sp.F16X3, seed=0)
    pts = torch.from_numpy(synth.lidar_cloud(cfg, args.frame)).to(dev)
    L = lib()
    L.p3d_debug_f16_timeline.argtypes = [C.c_void_p, C.c_int]
    bufs = []
    orig = L.p3d_sparse_conv_f16

    class Hook:
        def __call__(self, *a):
            buf = torch.zeros((8200,), dtype=torch.int64, device=dev)
            L.p3d_debug_f16_timeline(C.c_void_p(buf.data_ptr()), args.cta)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            rc = orig(*a)
            e.record()
            L.p3d_debug_f16_timeline(None, 0)
            bufs.append((a[5].value if hasattr(a[5], "value") else a[5], a[6], a[4], buf, s, e))
            return rc

    mean, coors, npv, nv = vox.voxelize_mean(pts, cfg["voxel_size"], cfg["point_cloud_range"], cfg["max_points"],
                                             cfg["max_voxels"], 0)
    # warm pass (allocations, weight packing), then the probed pass
    pipe.net(mean, coors, 1, num=nv)
    torch.cuda.synchronize()
    L.p3d_sparse_conv_f16 = Hook()
    try:
        pipe.net(mean, coors, 1, num=nv)
        torch.cuda.synchronize()
    finally:
        L.p3d_sparse_conv_f16 = orig
    for cin, cout, K, buf, s, e in bufs:
        d = buf.cpu().numpy()
        t0, t1, t2, n_work, splits = [int(x) for x in d[8192:8197]]
        uses = d[:4096].reshape(512, 8)
        items = d[4096:4096 + 512].reshape(64, 8)
        nu = int((uses[:, 5] != 0).sum())
        ni = int((items[:, 6] != 0).sum())
        rec = {"env": {k: os.environ.get(k) for k in ("P3D_F16_NPW", "P3D_F16_NSUB", "P3D_F16_FLAGS") if os.environ.get(k)},
               "cin": int(cin), "cout": int(cout), "K": int(K), "us": round(s.elapsed_time(e) * 1e3, 1),
               "n_work": n_work, "splits": splits, "cta_items": ni, "cta_stage_uses": nu,
               "total_cyc": t2 - t0, "prologue_cyc": t1 - t0}
        if nu > 1:
            u = uses[:nu].astype(np.int64)
            rec["prod_issue_cyc"] = med(u[:, 1] - u[:, 0])                    # gathers of one stage issued
            rec["prod_wait_empty_cyc"] = med(u[1:, 0] - u[:-1, 1])            # producer idle before next stage
            rec["gather_latency_cyc"] = med(u[:, 4] - u[:, 1])                # issued -> stage full at the MMA warp
            rec["w_issue_after_prod_cyc"] = med(u[:, 3] - u[:, 1])
            rec["mma_issue_cyc"] = med(u[:, 5] - u[:, 4])
            rec["mma_idle_cyc"] = med(u[1:, 4] - u[:-1, 5])                   # MMA warp waiting for the next full stage
            rec["stage_period_cyc"] = med(u[1:, 4] - u[:-1, 4])
            rec["first_full_after_prologue"] = int(u[0, 4] - t1)
        if ni > 0:
            it = items[:ni].astype(np.int64)
            rec["nbr_copy_cyc"] = med(it[:, 1] - it[:, 0])
            rec["prod_item_cyc"] = med(it[:, 4] - it[:, 3])
            rec["epi_cyc"] = med(it[:, 6] - it[:, 5])
            rec["item0"] = {"nbr_free": int(it[0, 0] - t0), "map": int(it[0, 1] - t0),
                            "prod_start": int(it[0, 3] - t0), "prod_end": int(it[0, 4] - t0),
                            "epi_start": int(it[0, 5] - t0), "epi_end": int(it[0, 6] - t0)}
            rec["last_item"] = {"prod_end": int(it[-1, 4] - t0), "epi_start": int(it[-1, 5] - t0),
                                "epi_end": int(it[-1, 6] - t0), "exit": t2 - t0}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
