#!/usr/bin/env python
"""In-kernel timeline + knock-out attribution for the split-layout tensor-core sparse conv (sparse_conv_tc2.cu).

Takes the neighbour maps of one real C3 frame (captured through sparse_nn.PROFILE), and for each distinct layer shape
launches the kernel through the debug entry point p3d_debug_split_launch with
  * flags 0           the real kernel (graph-replayed for the time, once more for CTA 0's clock64 timeline)
  * flags 2 / 4 / 1   without the gathers / the MMAs / the weight copies
  * flags 1|2|4       the bare skeleton (barriers, neighbour-map load, epilogue)
and prints one JSON line per (layer, splits). Needs a GPU:  python tools/split_probe.py [> profiles/...jsonl]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from paddle3d_b200 import synth  # noqa: E402
from paddle3d_b200._lib import lib  # noqa: E402
from paddle3d_b200._mem import ptr, stream  # noqa: E402
from paddle3d_b200.ops import sparse_nn as sp  # noqa: E402
from paddle3d_b200.ops import voxelize as vox  # noqa: E402
from paddle3d_b200.pipeline import CenterPointHotPath  # noqa: E402


def graph_time(fn, iters=20):
    st = torch.cuda.current_stream()
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    g.replay()
    b.record(st)
    b.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    L = lib()
    raw = L
    dbg_launch = raw.p3d_debug_split_launch
    dbg_launch.restype = ctypes.c_int
    vp, i64, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    dbg_launch.argtypes = [vp, vp, vp, i64, ci, ci, ci, vp, vp, vp, ci, vp]
    set_flags = raw.p3d_debug_set_flags
    set_flags.restype, set_flags.argtypes = ci, [ci]

    cfg = synth.C3
    pipe = CenterPointHotPath(device="cuda:0", precision=sp.TF32X3_SPLIT)
    pts = torch.from_numpy(synth.lidar_cloud(cfg, 0)).to(dev)
    P, V = cfg["max_points"], cfg["max_voxels"]
    sp.PROFILE = []
    mean, coors, npv, nv = vox.voxelize_mean(pts, cfg["voxel_size"], cfg["point_cloud_range"], P, V, 0)
    x, _ = pipe.net.forward_sparse(mean, coors, 1, num=nv)
    x.values()
    torch.cuda.synchronize()
    prof, sp.PROFILE = sp.PROFILE, None

    seen = set()
    for (cin, cout, K, prec, nbr, num, _s, _e, _kern) in prof:
        if cin < 16 or (cin, cout, K) in seen:
            continue
        seen.add((cin, cout, K))
        cap = nbr.shape[0]
        rows = int(num[0].item())
        n_in = int(nbr.max().item()) + 1
        xs = torch.randn((max(n_in, 1), 2 * cin), device=dev)
        w = torch.randn((K, cin, cout), device=dev) * 0.1
        packed = torch.empty((L.p3d_sparse_conv_packed_weight_bytes(K, cin, cout) // 4,), dtype=torch.float32, device=dev)
        assert L.p3d_sparse_conv_pack_weights(ptr(w), K, cin, cout, ptr(packed), stream(dev)) == 0
        for splits in (1, 2):
            out = torch.empty((splits * cap, cout), device=dev)
            dbg = torch.zeros((4096 + 260,), dtype=torch.int64, device=dev)
            rec = {"cin": cin, "cout": cout, "K": K, "rows": rows, "tiles": (rows + 127) // 128, "splits": splits}

            def launch(d):
                rc = dbg_launch(ptr(xs), ptr(nbr), ptr(num), cap, K, cin, cout, ptr(packed), ptr(out), ptr(d), splits,
                                stream(dev))
                assert rc == 0, rc

            for name, fl in (("full", 0), ("no_gather", 2), ("no_mma", 4), ("no_weights", 1), ("skeleton", 7),
                             ("skel_nofence", 7 | 8), ("skel_arrive_not_commit", 7 | 16),
                             ("full_nofence", 8)):
                assert set_flags(fl) == 0
                rec["us_" + name] = round(graph_time(lambda: launch(dbg)), 2)
            set_flags(0)
            dbg.zero_()
            launch(dbg)
            torch.cuda.synchronize()
            t = dbg.cpu().numpy()
            it = t[4096:4096 + 256].reshape(64, 4)
            k_entry, k_pro, k_exit = (int(v) for v in t[4096 + 256:4096 + 259])
            items = [r for r in it if r[0] > 0]
            rec["cta0"] = {
                "items": len(items),
                "prologue_cyc": k_pro - k_entry,
                "total_cyc": k_exit - k_entry,
                "nbr_load_cyc": [int(r[1] - r[0]) for r in items][:4],
                "mainloop_cyc": [int(r[2] - r[1]) for r in items][:4],
                "epilogue_cyc": [int(r[3] - r[2]) for r in items][:4],
            }
            u = t[:4096].reshape(512, 8)
            full = u[:, 5]
            n_u = int((full > 0).sum())
            if n_u > 4:
                d = np.diff(full[:n_u])
                rec["cta0"]["uses"] = n_u
                rec["cta0"]["cyc_per_use_median"] = int(np.median(d))
                rec["cta0"]["first_full_after_nbr"] = int(full[0] - items[0][1]) if items else None
                rec["cta0"]["mma_issue_cyc_median"] = int(np.median(u[:n_u, 6] - u[:n_u, 5]))
            if splits == 1:
                # raw window: per use (relative to use 8's slot-free time) the six recorded clocks, real and skeleton
                for tag, fl in (("real", 0), ("skel", 7)):
                    set_flags(fl)
                    dbg.zero_()
                    launch(dbg)
                    torch.cuda.synchronize()
                    uu = dbg.cpu().numpy()[:4096].reshape(512, 8)
                    base = int(uu[8, 0])
                    rec["window_" + tag] = [[int(uu[k, c] - base) if uu[k, c] else None for c in (0, 1, 2, 3, 5, 6)]
                                            for k in range(8, 20)]
                set_flags(0)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
