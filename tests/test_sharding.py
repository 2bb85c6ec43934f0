"""CPU suite: the N > 1 path (frame sharding + weight broadcast + max-over-ranks timing) on gloo, world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from paddle3d_b200 import synth
    from paddle3d_b200.layers import SparseResNet3D
    from paddle3d_b200.sharding import broadcast_weights, frames_for_rank, max_over_ranks
    net = SparseResNet3D(5, synth.C3["voxel_size"], synth.C3["point_cloud_range"]).init_weight(seed=100 + rank, device="cpu",
                                                                                           randomize_bn=True)
    before = float(net.conv_input[0].weight.sum())
    broadcast_weights(net, 0)
    ref = SparseResNet3D(5, synth.C3["voxel_size"], synth.C3["point_cloud_range"]).init_weight(seed=100, device="cpu",
                                                                                           randomize_bn=True)
    same = all(torch.equal(a.weight, b.weight) for a, b in zip(net.all_layers(), ref.all_layers()))
    folded = all(torch.equal(a._folded[0], b._folded[0]) for a, b in zip(net.all_layers(), ref.all_layers())
                 if hasattr(a, "_folded"))
    mx = max_over_ranks([1.0 + rank, 5.0 - rank], "cpu")
    out[rank] = dict(frames=frames_for_rank(11, rank, world), same=same, folded=folded, changed=(rank == 0) or before != float(ref.conv_input[0].weight.sum()), mx=mx)
    dist.destroy_process_group()


def test_frame_sharding_and_weight_broadcast_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    frames = sorted(f for r in range(world) for f in out[r]["frames"])
    assert frames == list(range(11))                       # every frame exactly once
    assert set(out[0]["frames"]).isdisjoint(out[1]["frames"])
    for r in range(world):
        assert out[r]["same"] and out[r]["folded"] and out[r]["changed"]
        assert out[r]["mx"] == [2.0, 5.0]                  # max over ranks


def test_frames_for_rank_edge_cases():
    from paddle3d_b200.sharding import frames_for_rank
    assert frames_for_rank(0, 0, 8) == []
    assert frames_for_rank(3, 5, 8) == []
    assert sum(len(frames_for_rank(1000, r, 8)) for r in range(8)) == 1000
