"""Frame-parallel multi-GPU plumbing: frames are independent units, so a sweep shards by frame
across one process per GPU.  The only collective is ONE broadcast of the weights at start-up (NCCL over
NVLink on GPUs, gloo in the CPU tests) plus the max-over-ranks reduction of the timing; there is no per-frame
exchange and therefore nothing to fuse with a kernel.  Reference counterpart: evaluation is per sample
(voxelizers/voxelize.py:68-73, postprocess.cu:19-20) and single-device (apis/trainer.py:49-51)."""
import torch
import torch.distributed as dist


def frames_for_rank(num_frames, rank, world):
    """Round-robin ownership: frame f belongs to rank f % world."""
    return list(range(rank, num_frames, world))


def broadcast_weights(net, src=0):
    """Broadcast every parameter / running statistic of a SparseResNet3D-like object (anything with all_layers())
    from `src`, then refresh the derived (folded / packed) copies."""
    from .ops import sparse_nn as sp
    for layer in net.all_layers():
        for name in ("weight", "bias", "_mean", "_variance"):
            t = getattr(layer, name, None)
            if isinstance(t, torch.Tensor):
                dist.broadcast(t, src)
    for layer in net.all_layers():
        if isinstance(layer, sp.BatchNorm):
            layer._bias_fold.clear()
            layer.set_parameters(layer.weight, layer.bias, layer._mean, layer._variance)
        elif hasattr(layer, "_packed"):
            layer._packed = None


def max_over_ranks(values, device):
    """Timing rule: every multi-GPU number is the max over ranks."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]
