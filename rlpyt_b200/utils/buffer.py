"""Buffer helpers (host-side mirror of ``rlpyt/utils/buffer.py``): build and move
tuple-of-array structures.  B200-specific additions: ``buffer_from_example`` can allocate
page-locked host memory (for async H2D of the step buffer) or device memory (the resident
``[T,B]`` sample buffers) instead of the reference's OS shared-memory numpy arrays.
"""
import multiprocessing as mp
import ctypes

import numpy as np
import torch

from rlpyt_b200.utils.collections import namedarraytuple_like


def _walk(buf, leaf):
    """Apply ``leaf`` to every array/tensor of a (named)tuple structure, keep ``None``."""
    if buf is None:
        return None
    if isinstance(buf, (np.ndarray, torch.Tensor)):
        return leaf(buf)
    parts = tuple(_walk(b, leaf) for b in buf)
    return parts if type(buf) is tuple else buf._make(parts)


def np_mp_array(shape, dtype):
    """numpy array on fork-shared memory (rlpyt/utils/buffer.py:55-62)."""
    shape = tuple(shape) if isinstance(shape, (list, tuple)) else (shape,)
    size = int(np.prod(shape))
    raw = mp.RawArray(ctypes.c_char, max(1, size * np.dtype(dtype).itemsize))
    return np.frombuffer(raw, dtype=dtype, count=size).reshape(shape)


def build_array(example, leading_dims, share_memory=False, where="numpy", device=None):
    """One leaf: ``leading_dims + example.shape`` zeros of example's dtype
    (rlpyt/utils/buffer.py:40-52).  ``where``: "numpy" | "pinned" | "cuda"."""
    if isinstance(example, torch.Tensor):
        example = example.detach().cpu().numpy()
    a = np.asarray(example)
    if a.dtype == "object":
        raise TypeError("Buffer example value cannot cast as np.dtype==object.")
    if not isinstance(leading_dims, (list, tuple)):
        leading_dims = (leading_dims,)
    shape = tuple(leading_dims) + a.shape
    if where == "numpy":
        return np_mp_array(shape, a.dtype) if share_memory else np.zeros(shape, dtype=a.dtype)
    tdtype = torch.from_numpy(np.zeros(1, dtype=a.dtype)).dtype
    if where == "pinned":
        t = torch.zeros(shape, dtype=tdtype)
        return t.pin_memory() if torch.cuda.is_available() else t
    if where == "cuda":
        return torch.zeros(shape, dtype=tdtype, device=device)
    raise ValueError(where)


def buffer_from_example(example, leading_dims, share_memory=False, where="numpy", device=None):
    """Allocate a buffer with the structure of ``example`` and extra leading dims
    (rlpyt/utils/buffer.py:11-37)."""
    if example is None:
        return None
    try:
        cls = namedarraytuple_like(example)
    except TypeError:
        return build_array(example, leading_dims, share_memory, where, device)
    return cls(*(buffer_from_example(v, leading_dims, share_memory, where, device) for v in example))


def torchify_buffer(buffer_):
    """numpy leaves -> torch tensors sharing memory (rlpyt/utils/buffer.py:120-135)."""
    return _walk(buffer_, lambda x: torch.from_numpy(x) if isinstance(x, np.ndarray) else x)


def numpify_buffer(buffer_):
    """torch leaves -> numpy (device tensors are copied to host) (buffer.py:138-153)."""
    return _walk(buffer_, lambda x: x.cpu().numpy() if isinstance(x, torch.Tensor) else x)


def buffer_to(buffer_, device=None, non_blocking=False):
    """Move every tensor leaf (rlpyt/utils/buffer.py:156-170)."""
    def leaf(x):
        if isinstance(x, np.ndarray):
            raise TypeError("Cannot move numpy array to device.")
        return x.to(device, non_blocking=non_blocking)
    return _walk(buffer_, leaf)


def buffer_method(buffer_, method_name, *args, **kwargs):
    """Call a method on every leaf (rlpyt/utils/buffer.py:173-187)."""
    return _walk(buffer_, lambda x: getattr(x, method_name)(*args, **kwargs))


def buffer_func(buffer_, func, *args, **kwargs):
    """Call ``func(leaf, ...)`` on every leaf (rlpyt/utils/buffer.py:190-205)."""
    return _walk(buffer_, lambda x: func(x, *args, **kwargs))


def get_leading_dims(buffer_, n_dim=1):
    """Common leading dims of all leaves (rlpyt/utils/buffer.py:208-220)."""
    if buffer_ is None:
        return None
    if isinstance(buffer_, (np.ndarray, torch.Tensor)):
        return tuple(buffer_.shape[:n_dim])
    found = {get_leading_dims(b, n_dim) for b in buffer_ if b is not None}
    if len(found) != 1:
        raise ValueError(f"Found mismatched leading dimensions: {found}")
    return found.pop()
