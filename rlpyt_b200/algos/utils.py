"""Drop-in for ``rlpyt/algos/utils.py`` backed by the sm_100a kernels (csrc/returns.cu).

Same names, argument meaning and return conventions as the reference functions
(rlpyt/algos/utils.py:8, :24, :67, :104): time-major ``[T, ...]`` inputs, optional
``*_dest`` buffers that are written in place and returned.  Inputs may be

* CUDA tensors  - the fast path: kernels run on the tensors' storage, outputs are CUDA
  tensors (device-resident, nothing crosses PCIe);
* CPU tensors / numpy arrays - the host-buffer path: staged to the GPU, computed there and
  copied back into the same container type (this is H2D + kernel + D2H, not a CPU
  implementation; there is none in this package).

``algo``: 0 auto, 1 streaming column kernel (bit-identical to the reference's fp32
operation order), 2 T-parallel warp-segmented scan (<=1e-5 relative).
"""
import numpy as np
import torch

from rlpyt_b200 import _lib


def _device():
    if not torch.cuda.is_available():
        raise _lib.B200LibraryError("rlpyt_b200 needs a CUDA device (B200); no CPU fallback exists")
    return torch.device("cuda", torch.cuda.current_device())


def _stage(x, dtype):
    """-> contiguous CUDA tensor of ``dtype`` (no copy if already so)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if not x.is_cuda:
        x = x.to(_device(), non_blocking=False)
    if x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


def _done_u8(done):
    """``done`` (bool / uint8 / the reference's float cast, numpy or torch) -> CUDA uint8 0/1."""
    if isinstance(done, np.ndarray):
        done = torch.from_numpy(np.ascontiguousarray(done))
    if not isinstance(done, torch.Tensor):
        done = torch.as_tensor(done)
    if not done.is_cuda:
        done = done.to(_device())
    if done.dtype == torch.bool:
        return done.contiguous().view(torch.uint8)
    if done.dtype == torch.uint8:
        return done.contiguous()
    return (done != 0).contiguous().view(torch.uint8)


def _deliver(dev, like, dest):
    """Return ``dev`` (CUDA result) in the container kind of ``like``; fill ``dest`` if given."""
    if dest is not None:
        if isinstance(dest, np.ndarray):
            dest[...] = dev.cpu().numpy().reshape(dest.shape)
        elif dest.data_ptr() != dev.data_ptr():
            dest.copy_(dev.reshape(dest.shape))
        return dest
    if isinstance(like, np.ndarray):
        return dev.cpu().numpy()
    if isinstance(like, torch.Tensor) and not like.is_cuda:
        return dev.cpu()
    return dev


def _tb(x):
    T = x.shape[0]
    B = int(x.numel() // T) if T > 0 else 0
    return T, B


def _out_buf(dest, shape, dtype, device):
    """Write straight into ``dest`` when it is a matching contiguous CUDA tensor."""
    if (isinstance(dest, torch.Tensor) and dest.is_cuda and dest.dtype == dtype
            and dest.is_contiguous() and tuple(dest.shape) == tuple(shape)):
        return dest
    return torch.empty(shape, dtype=dtype, device=device)


def discount_return(reward, done, bootstrap_value, discount, return_dest=None, algo=0):
    """rlpyt/algos/utils.py:8-21."""
    r = _stage(reward, torch.float32)
    d = _done_u8(done)
    bv = _stage(bootstrap_value, torch.float32).reshape(-1)
    T, B = _tb(r)
    assert d.numel() == r.numel() and bv.numel() == B
    ret = _out_buf(return_dest, r.shape, torch.float32, r.device)
    with torch.cuda.device(r.device):
        _lib.call("rl_discount_return_f32", _lib.ptr(r), _lib.ptr(d), _lib.ptr(bv), None,
                  _lib.ptr(ret), None, T, B, float(discount), int(algo), _lib.stream())
    return _deliver(ret, reward, return_dest)


def generalized_advantage_estimation(reward, value, done, bootstrap_value, discount, gae_lambda,
                                     advantage_dest=None, return_dest=None, algo=0):
    """rlpyt/algos/utils.py:24-40.  Returns ``(advantage, return_)``."""
    r = _stage(reward, torch.float32)
    v = _stage(value, torch.float32)
    d = _done_u8(done)
    bv = _stage(bootstrap_value, torch.float32).reshape(-1)
    T, B = _tb(r)
    assert v.numel() == r.numel() == d.numel() and bv.numel() == B
    adv = _out_buf(advantage_dest, r.shape, torch.float32, r.device)
    ret = _out_buf(return_dest, r.shape, torch.float32, r.device)
    gl = float(np.float32(float(discount) * float(gae_lambda)))  # utils.py:38 (python double product)
    with torch.cuda.device(r.device):
        _lib.call("rl_gae_f32", _lib.ptr(r), _lib.ptr(v), _lib.ptr(d), _lib.ptr(bv), _lib.ptr(adv),
                  _lib.ptr(ret), T, B, float(discount), gl, int(algo), _lib.stream())
    return _deliver(adv, reward, advantage_dest), _deliver(ret, reward, return_dest)


_GPOW_CACHE = {}


def _discount_pow(discount, n_step, device):
    key = (float(discount), int(n_step), str(device))
    t = _GPOW_CACHE.get(key)
    if t is None:
        t = torch.tensor([float(discount) ** k for k in range(n_step)], dtype=torch.float32, device=device)
        _GPOW_CACHE[key] = t
    return t


def discount_return_n_step(reward, done, n_step, discount, return_dest=None, done_n_dest=None,
                           do_truncated=False):
    """rlpyt/algos/utils.py:67-101.  Returns ``(return_, done_n)``; ``done_n`` has the dtype of
    ``done`` (bool stays bool)."""
    r = _stage(reward, torch.float32)
    d = _done_u8(done)
    T_in, B = _tb(r)
    rlen = T_in if do_truncated else T_in - (n_step - 1)
    shape = (rlen,) + tuple(r.shape[1:])
    ret = _out_buf(return_dest, shape, torch.float32, r.device)
    dn = torch.empty(shape, dtype=torch.uint8, device=r.device)
    with torch.cuda.device(r.device):
        _lib.call("rl_nstep_return_f32", _lib.ptr(r), _lib.ptr(d), _lib.ptr(_discount_pow(discount, n_step, r.device)),
                  _lib.ptr(ret), _lib.ptr(dn), T_in, B, int(n_step), int(bool(do_truncated)), _lib.stream())
    done_dtype = done.dtype if isinstance(done, (torch.Tensor, np.ndarray)) else torch.bool
    if isinstance(done, np.ndarray):
        dn_out = dn.cpu().numpy().astype(done_dtype)
        if done_n_dest is not None:
            done_n_dest[...] = dn_out
            dn_out = done_n_dest
    else:
        dn_t = dn.view(torch.bool) if done_dtype == torch.bool else dn.to(done_dtype)
        if not (isinstance(done, torch.Tensor) and done.is_cuda):
            dn_t = dn_t.cpu()
        if done_n_dest is not None:
            done_n_dest.copy_(dn_t)
            dn_t = done_n_dest
        dn_out = dn_t
    return _deliver(ret, reward, return_dest), dn_out


def valid_from_done(done):
    """rlpyt/algos/utils.py:104-112.  Float32 mask, zero after the first ``done`` of a column."""
    d = _done_u8(done)
    T, B = _tb(d)
    valid = torch.empty(d.shape, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.call("rl_valid_from_done_f32", _lib.ptr(d), _lib.ptr(valid), T, B, _lib.stream())
    return _deliver(valid, done, None)


_NORM_SCRATCH = {}


def normalize_advantage_(advantage, valid=None, stats_out=None):
    """In-place ``(adv - mean) / max(std, 1e-6)`` over ``valid > 0`` - the normalisation step of
    ``process_returns`` (rlpyt/algos/pg/base.py:65-73).  CUDA tensors only."""
    _lib.require_cuda(advantage, valid)
    assert advantage.dtype == torch.float32 and advantage.is_contiguous()
    n = advantage.numel()
    lib = _lib.load()
    nbytes = int(lib.rl_adv_normalize_scratch_bytes(n))
    key = (str(advantage.device), nbytes)
    scratch = _NORM_SCRATCH.get(key)
    if scratch is None:
        scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=advantage.device)
        _NORM_SCRATCH[key] = scratch
    if valid is not None:
        assert valid.dtype == torch.float32 and valid.is_contiguous() and valid.numel() == n
    with torch.cuda.device(advantage.device):
        _lib.call("rl_adv_normalize_f32", _lib.ptr(advantage), _lib.ptr(valid), n, _lib.ptr(scratch),
                  _lib.ptr(stats_out), _lib.stream(), n_launch=2)
    return advantage
