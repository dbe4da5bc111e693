#!/usr/bin/env python
"""Per-kernel micro-benchmarks (CUDA events, warm-up, L2 flush between reps).  Prints JSON
lines; used to fill DESIGN.md's roofline table and to pick dispatch crossovers.

    python tools/microbench.py returns [--reps 20]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                        "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = float(PEAKS.get("hbm_gbs", 6650.0))

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    _flush.fill_(1)


def timeit(fn, reps=20, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def out(**kw):
    print(json.dumps(kw), flush=True)


def bench_returns(reps):
    from rlpyt_b200.algos import utils as U
    for (T, B) in [(128, 256), (128, 4096), (128, 1 << 14), (128, 1 << 16), (128, 1 << 18), (128, 1 << 20)]:
        gen = torch.Generator(device="cuda").manual_seed(0)
        r = torch.randn(T, B, device="cuda", generator=gen)
        v = torch.randn(T, B, device="cuda", generator=gen)
        d = (torch.rand(T, B, device="cuda", generator=gen) < 0.01)
        b = torch.randn(1, B, device="cuda", generator=gen)
        adv = torch.empty_like(r)
        ret = torch.empty_like(r)
        for algo in (1, 2):
            if algo == 2 and B > (1 << 16):
                continue
            for flush in (True, False):
                med, best = timeit(lambda: U.generalized_advantage_estimation(
                    r, v, d, b, 0.99, 0.98, advantage_dest=adv, return_dest=ret, algo=algo), reps, flush=flush)
                nbytes = T * B * 17 + 4 * B
                out(kernel="gae", algo=algo, T=T, B=B, l2_flush=flush, us_med=med * 1e6, us_best=best * 1e6,
                    GBs=nbytes / med / 1e9, frac_hbm=nbytes / med / 1e9 / HBM)
            med, best = timeit(lambda: U.discount_return(r, d, b, 0.99, return_dest=ret, algo=algo), reps)
            nbytes = T * B * 9 + 4 * B
            out(kernel="discount_return", algo=algo, T=T, B=B, l2_flush=True, us_med=med * 1e6,
                GBs=nbytes / med / 1e9, frac_hbm=nbytes / med / 1e9 / HBM)
    # reference-style copy for context: same bytes through torch's copy kernel
    n = 128 * (1 << 20)
    a = torch.empty(n, dtype=torch.float32, device="cuda")
    c = torch.empty_like(a)
    med, best = timeit(lambda: c.copy_(a), reps)
    out(kernel="torch_copy_f32", n=n, us_med=med * 1e6, GBs=2 * 4 * n / med / 1e9)


def bench_gemm(reps):
    from rlpyt_b200.models.gemm_op import gemm_tn
    for name, (M, N, K) in dict(fc_fwd=(8192, 512, 3200), fc_dgrad=(8192, 3200, 512), fc_wgrad=(512, 3200, 8192),
                                fc_step=(256, 512, 3200)).items():
        a = torch.randn(M, K, device="cuda")
        b = torch.randn(N, K, device="cuda")
        med, best = timeit(lambda: gemm_tn(a, b), reps, flush=False)
        med_t, _ = timeit(lambda: torch.mm(a, b.t()), reps, flush=False)
        fl = 2.0 * M * N * K
        out(kernel="gemm_tf32x3", shape=name, M=M, N=N, K=K, us_med=med * 1e6, us_best=best * 1e6,
            eff_TFLOPs=fl / med / 1e12, tensor_TFLOPs_issued=3 * fl / med / 1e12, torch_fp32_us=med_t * 1e6,
            speedup_vs_cublas_fp32=med_t / med)


BENCHES = {"returns": bench_returns, "gemm": bench_gemm}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=list(BENCHES))
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    out(device=torch.cuda.get_device_name(0), hbm_peak_gbs=HBM)
    for w in a.which:
        BENCHES[w](a.reps)
