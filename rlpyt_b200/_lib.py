"""ctypes binding of librlpyt_b200.so (the C ABI declared in include/rlpyt_b200.h).

The product path has NO CPU fallback: if the library is missing or a call fails, this
module raises.  Device memory, streams and collectives come from torch (plumbing); the
arithmetic runs in the hand-written sm_100a kernels behind these entry points.
"""
import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p, c_double

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librlpyt_b200.so")

P = c_void_p  # every device / host pointer crosses as void*

# name -> (restype, [argtypes]) ; mirrors include/rlpyt_b200.h one to one.
SIGNATURES = {
    "rl_b200_abi_version": (c_int, []),
    "rl_b200_last_error": (c_char_p, []),
    "rl_b200_sm_count": (c_int, []),
    "rl_host_stream_copy": (c_int, [P, P, c_int64]),
    "rl_upload_async": (c_int, [P, P, c_int64, P]),
    "rl_gae_f32": (c_int, [P, P, P, P, P, P, c_int, c_int64, c_float, c_float, c_int, P]),
    "rl_discount_return_f32": (c_int, [P, P, P, P, P, P, c_int, c_int64, c_float, c_int, P]),
    "rl_nstep_return_f32": (c_int, [P, P, P, P, P, c_int, c_int64, c_int, c_int, P]),
    "rl_valid_from_done_f32": (c_int, [P, P, c_int, c_int64, P]),
    "rl_adv_normalize_scratch_bytes": (c_int64, [c_int64]),
    "rl_adv_normalize_f32": (c_int, [P, P, c_int64, P, P, P]),
    "rl_pg_loss_scratch_bytes": (c_int64, [c_int64]),
    "rl_ppo_loss_f32": (c_int, [P, P, P, P, P, P, P, c_int64, c_int, c_float, c_float, c_float, P, P, P, P, P]),
    "rl_ppo_loss_devclip_f32": (c_int, [P, P, P, P, P, P, P, c_int64, c_int, P, c_float, c_float, P, P, P, P, P]),
    "rl_pg_heads_forward_f32": (c_int, [P, P, P, P, P, P, P, c_int64, c_int, c_int, P]),
    "rl_pg_heads_backward_scratch_bytes": (c_int64, [c_int64, c_int, c_int]),
    "rl_pg_heads_backward_f32": (c_int, [P, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, P, P]),
    "rl_gather_rows": (c_int, [P, P, P, c_int64, c_int64, P]),
    "rl_gather_rows_multi": (c_int, [c_int, P, P, P, P, c_int64, P]),
    "rl_clip_adam_scratch_bytes": (c_int64, [c_int64]),
    "rl_clip_adam_f32": (c_int, [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int64,
                                 c_float, c_float, P, P, P]),
    "rl_sumtree_find_f64": (c_int, [P, c_int, P, c_int64, c_int64, P, P, P, P, P, P]),
    "rl_sumtree_update_f64": (c_int, [P, c_int, P, c_int64, P, c_double, c_int64, P, P]),
    "rl_sumtree_update_batch": (c_int, [P, c_int, P, P, c_float, P, c_int64, P, P, P]),
    "rl_pow_f32_to_f64": (c_int, [P, c_float, P, c_int64, P]),
    "rl_is_weights_f32": (c_int, [P, c_double, P, c_int, P]),
    "rl_is_weights_eps_f32": (c_int, [P, c_double, c_double, P, c_int, P]),
    "rl_replay_extract_sequences": (c_int, [P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, P, P, c_int64,
                                            c_int64, P, P, P, P, P, P, P]),
    "rl_replay_extract": (c_int, [P, P, P, P, P, P, c_int64, c_int64, c_int64, c_int, c_int, P, P, c_int64,
                                  P, P, P, P, P, P, P, P, P, P, P]),
    "rl_conv1_u8_forward": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P]),
    "rl_conv1_u8_wgrad_scratch_bytes": (c_int64, []),
    "rl_conv1_u8_wgrad": (c_int, [P, P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P, P]),
    "rl_gemm_tf32x3_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "rl_gemm_tf32x3_f32": (c_int, [P, P, P, P, c_int64, c_int64, c_int64, c_int, P, P]),
    "rl_gemm_ts_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "rl_gemm_ts_f32": (c_int, [P, c_int, P, P, P, P, c_int, c_int64, c_int64, c_int64, c_int, P, P]),
    "rl_gemm_ts_masked_f32": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int64, P, P]),
    "rl_split_lo_f32": (c_int, [P, P, c_int64, P]),
    "rl_transpose_split_f32": (c_int, [P, P, P, c_int64, c_int64, P]),
    "rl_conv1_u8_forward_tc": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P]),
    "rl_conv2_forward_tc": (c_int, [P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P]),
    "rl_conv2_dgrad_tc_scratch_bytes": (c_int64, []),
    "rl_conv2_dgrad_tc": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "rl_dqn_loss_scratch_bytes": (c_int64, [c_int64]),
    "rl_dqn_loss_f32": (c_int, [P, P, P, P, P, P, P, c_int64, c_int, c_float, c_float, P, P, P, P, P]),
    "rl_relu_backward_f32": (c_int, [P, P, P, c_int64, P]),
    "rl_transpose_f32": (c_int, [P, P, c_int64, c_int64, P]),
    "rl_conv_wgrad_tc_scratch_bytes": (c_int64, []),
    "rl_conv1_u8_wgrad_tc": (c_int, [P, P, P, P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "rl_conv2_wgrad_tc": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "rl_conv1_u8_i8_supported": (c_int, [c_int, c_int, c_int]),
    "rl_conv1_u8_forward_i8": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P]),
    "rl_conv1_u8_forward_i8_stream": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P]),
    "rl_conv1_u8_wgrad_i8_scratch_bytes": (c_int64, []),
    "rl_conv1_u8_wgrad_i8": (c_int, [P, P, P, P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "rl_conv1_u8_wgrad_i8_scaled": (c_int, [P, P, P, P, P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "rl_conv2_s2d_supported": (c_int, [c_int, c_int, c_int]),
    "rl_conv2_forward_s2d": (c_int, [P, P, P, P, c_int64, c_int, c_int, c_int, c_int, P]),
    "rl_conv2_dgrad_s2d": (c_int, [P, P, P, c_int64, c_int, c_int, c_int, P]),
    "rl_conv2_dgrad_s2d_absmax": (c_int, [P, P, P, P, c_int64, c_int, c_int, c_int, P]),
    "rl_categorical_sample_f32": (c_int, [P, P, P, P, P, c_int64, c_int, P]),
    "rl_conv2_wgrad_s2d_scratch_bytes": (c_int64, []),
    "rl_conv2_wgrad_s2d": (c_int, [P, P, P, P, c_int64, c_int, c_int, c_int, P, P]),
    "rl_pg_head_sample_f32": (c_int, [P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_int, P]),
    "rl_a2c_loss_f32": (c_int, [P, P, P, P, P, P, c_int64, c_int, c_float, c_float, P, P, P, P, P]),
}

_lock = threading.Lock()
_lib = None
launch_count = 0  # kernels launched through this binding (bench.py reports it)


class B200LibraryError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200LibraryError(
                f"{LIB_PATH} not found: build it with `python -m rlpyt_b200.csrc.build` "
                "(or __graft_entry__.build()).  rlpyt_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().rl_b200_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise B200LibraryError(f"{what} failed (rc={rc}): {last_error()}")


def ptr(t):
    """Device (or pinned-host) pointer of a contiguous tensor, or NULL for None."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("rlpyt_b200 kernels need contiguous tensors")
    return c_void_p(t.data_ptr())


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B200LibraryError(
                "rlpyt_b200 kernels run on CUDA tensors only (got a CPU tensor); "
                "there is no CPU fallback - move the data to the GPU first.")


def call(name, *args, n_launch=1):
    global launch_count
    rc = getattr(load(), name)(*args)
    check(rc, name)
    launch_count += n_launch


_c_char = ctypes.c_char


def host_stream_copy_ptr(dst_ptr, nbytes, dtype, src):
    """The same copy for a destination whose address the caller computed ONCE (a row of the step buffer): per call only
    the source's address is taken - through the buffer protocol (0.5 us) rather than ``ndarray.ctypes`` (1.3 us) - and
    its size / dtype compared.  Returns False (nothing copied) when the source is not a writable C-contiguous array of
    that size and type; the caller then takes ``host_stream_copy``.  The worker loop copies one 28 KB observation per
    environment step: the wrapper's bookkeeping was two thirds of the copy's cost."""
    try:
        if src.nbytes != nbytes or src.dtype != dtype:
            return False
        addr = ctypes.addressof(_c_char.from_buffer(src))
    except (TypeError, ValueError, BufferError, AttributeError):
        return False
    (_lib or load()).rl_host_stream_copy(dst_ptr, addr, nbytes)
    return True


def host_stream_copy(dst, src):
    """numpy -> numpy copy with non-temporal stores (see rl_host_stream_copy).  Falls back to a
    plain assignment for non-contiguous / mismatched arrays."""
    if (dst.flags["C_CONTIGUOUS"] and getattr(src, "flags", None) is not None and src.flags["C_CONTIGUOUS"]
            and dst.nbytes == src.nbytes and dst.dtype == src.dtype):
        load().rl_host_stream_copy(dst.ctypes.data, src.ctypes.data, dst.nbytes)
    else:
        dst[...] = src
