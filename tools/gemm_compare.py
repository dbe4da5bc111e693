#!/usr/bin/env python
"""The two fc GEMM kernels side by side (CUDA events, warm, L2 flushed between iterations by the operand sizes):
``python tools/gemm_compare.py`` prints one JSON line per shape with us per call for "ss" (csrc/gemm_tf32x3.cu)
and "ts" (csrc/gemm_ts.cuh, incl. its operand preparation), and for the whole Linear forward+backward."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlpyt_b200.models import gemm_op  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for M in (256, 512, 1024, 8192):
        x = torch.randn(M, 3200, device="cuda", generator=g)
        w = torch.randn(512, 3200, device="cuda", generator=g) / 56
        b = torch.randn(512, device="cuda", generator=g)
        w_lo = gemm_op.split_lo(w)
        out = {"shape": [M, 512, 3200],
               "ss_us": timed(lambda: gemm_op.gemm_tn(x, w, b, True)),
               "ts_us_presplit": timed(lambda: gemm_op.gemm_ts(x, w, w_lo, b, True)),
               "ts_us_incl_split": timed(lambda: gemm_op.gemm_ts(x, w, gemm_op.split_lo(w), b, True))}
        print(json.dumps(out), flush=True)
    M = 8192
    x = torch.randn(M, 3200, device="cuda", generator=g)
    w = (torch.randn(512, 3200, device="cuda", generator=g) / 56).requires_grad_(True)
    b = torch.randn(512, device="cuda", generator=g).requires_grad_(True)
    go = torch.randn(M, 512, device="cuda", generator=g)
    xr = x.clone().requires_grad_(True)

    def fwd_bwd():
        y = gemm_op.linear_tf32x3(xr, w, b, relu=True)
        y.backward(go)
        xr.grad = w.grad = b.grad = None

    res = {}
    for impl in ("ss", "ts"):
        gemm_op.GEMM_IMPL = impl
        res[impl + "_linear_fwd_bwd_us"] = timed(fwd_bwd)
    print(json.dumps({"shape": [M, 512, 3200], **res}), flush=True)


if __name__ == "__main__":
    main()
