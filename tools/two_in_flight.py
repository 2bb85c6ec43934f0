#!/usr/bin/env python
"""Experiment: throughput of the captured CenterPoint frame with ONE vs TWO frames in flight on one GPU (two pipelines with
their own buffers / graphs / streams, replays alternating).  The small latency-bound kernels of one frame (voxelize,
rulebooks, post-processing, conv tails) can fill SMs the other frame's kernels leave idle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from paddle3d_b200 import synth  # noqa: E402
from paddle3d_b200.ops import sparse_nn as sp  # noqa: E402
from paddle3d_b200.pipeline import CenterPointHotPath  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    cfg = synth.C3
    frames = bench.frame_pool(cfg, 16)
    dev_frames = [torch.from_numpy(f).to(dev) for f in frames]
    pipes = []
    for k in range(2):
        p = CenterPointHotPath(cfg, dev, precision=sp.F16X3, seed=0, with_head=True, keep_bev=False, bn_gain=bench.BN_GAIN)
        p.calibrate_head(dev_frames[0])
        p.points.copy_(dev_frames[0])
        p.capture()
        pipes.append(p)
    steps = 200

    def run(n_pipes):
        for w in range(6):
            p = pipes[w % n_pipes]
            with torch.cuda.stream(p.stream):
                p.points.copy_(dev_frames[w % 16], non_blocking=True)
                p.graph.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        base = torch.cuda.current_stream(dev)
        s.record(base)
        for p in pipes[:n_pipes]:
            p.stream.wait_event(s)
        for i in range(steps):
            p = pipes[i % n_pipes]
            with torch.cuda.stream(p.stream):
                p.points.copy_(dev_frames[i % 16], non_blocking=True)
                p.graph.replay()
        for p in pipes[:n_pipes]:
            base.wait_stream(p.stream)
        e.record(base)
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        return steps / (ms * 1e-3), ms / steps

    out = {}
    for n in (1, 2, 1, 2):
        fps, ms = run(n)
        out.setdefault("in_flight_%d" % n, []).append(round(fps, 1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
