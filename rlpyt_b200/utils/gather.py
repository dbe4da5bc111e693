"""Python face of the indexed row-gather kernels (csrc/gather.cu)."""
import ctypes

import torch

from rlpyt_b200 import _lib


def gather_rows(src, idx, out=None):
    """``src.view(R, -1)[idx]`` for a contiguous CUDA tensor ``src`` whose first dim is the row
    dim; ``idx`` int64 CUDA.  Returns ``[len(idx), *src.shape[1:]]``."""
    _lib.require_cuda(src, idx)
    assert src.is_contiguous() and idx.dtype == torch.int64 and idx.is_contiguous()
    n = idx.numel()
    row_bytes = (src.numel() // max(1, src.shape[0])) * src.element_size()
    if out is None:
        out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    if n == 0:
        return out
    with torch.cuda.device(src.device):
        _lib.call("rl_gather_rows", _lib.ptr(src), _lib.ptr(idx), _lib.ptr(out), n, row_bytes, _lib.stream())
    return out


def gather_rows_multi(srcs, idx, outs=None):
    """Gather several small fields (row size a multiple of 4 bytes; bool/uint8 fields are not
    eligible) with one index vector in ONE launch.  ``srcs``: list of contiguous CUDA tensors
    sharing the leading (row) dim."""
    n = idx.numel()
    k = len(srcs)
    assert 1 <= k <= 8
    if outs is None:
        outs = [torch.empty((n,) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device) for s in srcs]
    src_arr = (ctypes.c_void_p * k)(*[s.data_ptr() for s in srcs])
    dst_arr = (ctypes.c_void_p * k)(*[o.data_ptr() for o in outs])
    rb_arr = (ctypes.c_int64 * k)(*[(s.numel() // max(1, s.shape[0])) * s.element_size() for s in srcs])
    for s in srcs:
        _lib.require_cuda(s)
        assert s.is_contiguous()
    with torch.cuda.device(idx.device):
        _lib.call("rl_gather_rows_multi", k, src_arr, dst_arr, rb_arr, _lib.ptr(idx), n, _lib.stream())
    return outs


class LazyRows:
    """``src[rows]`` that has not been gathered yet: lets a consumer that can read rows in place
    (the fused first-layer kernel reads observations straight from the resident ``[T*B]`` batch)
    skip materialising the 231 MB minibatch copy.  ``materialize()`` gives the plain tensor."""

    def __init__(self, src, rows):
        _lib.require_cuda(src, rows)
        assert src.is_contiguous() and rows.dtype == torch.int64 and rows.is_contiguous()
        self.src, self.rows = src, rows

    @property
    def shape(self):
        return (self.rows.numel(),) + tuple(self.src.shape[1:])

    @property
    def dtype(self):
        return self.src.dtype

    @property
    def device(self):
        return self.src.device

    def dim(self):
        return self.src.dim()

    def materialize(self):
        return gather_rows(self.src, self.rows)


class HostMappedFrames:
    """uint8 frames ``[B,C,H,W]`` lying in PAGE-LOCKED host memory that the device can address (``cudaHostRegister`` /
    pinned allocations under unified addressing: the host pointer is the device pointer), together with the HBM tensor
    they are to be recorded in.  A first layer that streams its frames with bulk copies can read them from there and
    write the HBM copy in the same pass (csrc/conv1_i8.cuh ``copy_out``): the sampler's per-step H2D disappears from in
    front of ``agent.step``.  Quacks enough like a tensor for the agents' ``step`` plumbing."""

    is_cuda = True

    def __init__(self, host_ptr, copy_to):
        _lib.require_cuda(copy_to)
        assert copy_to.dtype == torch.uint8 and copy_to.is_contiguous() and copy_to.dim() == 4
        self.host_ptr, self.copy_to = int(host_ptr), copy_to

    @property
    def shape(self):
        return tuple(self.copy_to.shape)

    @property
    def dtype(self):
        return torch.uint8

    @property
    def device(self):
        return self.copy_to.device

    def dim(self):
        return 4
