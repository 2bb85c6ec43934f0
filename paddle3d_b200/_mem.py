"""Device memory / stream plumbing (torch is the allocator; the product is the CUDA library)."""
import ctypes as C

import torch

from ._lib import P3DError

_WS = {}
_RETIRED = []  # outgrown workspaces stay allocated: graphs captured earlier hold their addresses


def require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor):
        raise P3DError("%s must be a torch.Tensor on a CUDA device" % name)
    if not t.is_cuda:
        raise P3DError("%s must be a GPU tensor (paddle3d_b200 has no CPU path)" % name)
    if dtype is not None and t.dtype != dtype:
        raise P3DError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def workspace(nbytes, device, tag="default", zero=False):
    """Grow-only scratch buffer per (device, stream, tag); 256-byte aligned by the caching allocator.  A buffer that
    is replaced by a larger one is kept alive (a captured CUDA graph may still hold its address).  zero=True:
    zero-filled when created (for kernels that keep a self-cleaning region in it)."""
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _RETIRED.append(buf)
        alloc = torch.zeros if zero else torch.empty
        buf = alloc(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf
