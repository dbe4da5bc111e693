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


def make_replay(size, B=256, device_buf=True):
    from rlpyt_b200.replays.non_sequence.frame import PrioritizedReplayFrameBuffer
    from rlpyt_b200.utils.collections import namedarraytuple
    Ex = namedarraytuple("SamplesToBuffer", ["observation", "action", "reward", "done"])
    ex = Ex(observation=np.zeros((4, 84, 84), np.uint8), action=np.int64(0), reward=np.float32(0), done=np.bool_(False))
    buf = PrioritizedReplayFrameBuffer(example=ex, size=size, B=B, discount=0.99, n_step_return=3, alpha=0.6, beta=0.4,
                                       default_priority=1)
    T = 128
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    n_app = buf.T // T + 2
    for i in range(n_app):
        buf.append_samples(Ex(observation=obs, action=torch.randint(0, 6, (T, B), device="cuda", generator=g),
                              reward=torch.randn(T, B, device="cuda", generator=g),
                              done=torch.rand(T, B, device="cuda", generator=g) < 0.005))
    return buf, Ex


def bench_replay(reps):
    """BASELINE.json config 4: 1M-frame prioritized frame replay, batch 512, n-step 3."""
    import time
    buf, Ex = make_replay(1_000_000)
    np.random.seed(0)
    pri = torch.rand(512, device="cuda") + 0.01
    med, best = timeit(lambda: buf.sample_batch(512), reps, flush=True)
    out(kernel="replay.sample_batch(512)", frames=buf.size, us_med=med * 1e6, us_best=best * 1e6,
        samples_per_s=512 / med)
    buf.sample_batch(512)
    med_u, _ = timeit(lambda: buf.update_batch_priorities(pri), reps, flush=False)
    out(kernel="replay.update_batch_priorities(512)", us_med=med_u * 1e6)
    # extraction kernel alone (fixed indices): HBM-bound gather
    (T_idxs, B_idxs), _p = buf.priority_tree.sample(512)
    med_e, best_e = timeit(lambda: buf.extract_batch(T_idxs, B_idxs), reps, flush=True)
    nbytes = 2 * 2 * 512 * 4 * 84 * 84
    out(kernel="replay_extract_kernel", n=512, us_med=med_e * 1e6, us_best=best_e * 1e6, GBs=nbytes / med_e / 1e9,
        frac_hbm=nbytes / med_e / 1e9 / HBM, algorithmic_bytes=nbytes)
    t_adv, _ = timeit(lambda: buf.priority_tree.advance(0) or buf.priority_tree._update_segment(
        128 * 256, leaf_base=buf.priority_tree.low_idx, scalar=1.0), 5, flush=False)
    out(kernel="sumtree advance segment (32768 leaves, 20 levels)", us_med=t_adv * 1e6)
    # CPU oracle (reference algorithm restated) on a smaller ring: per-batch cost is size independent
    from oracle.replay import FrameReplay
    o = FrameReplay((4, 84, 84), 100_000, 256, discount=0.99, n_step_return=3)
    rng = np.random.default_rng(0)
    obs = rng.integers(0, 256, size=(64, 256, 4, 84, 84), dtype=np.uint8)
    for i in range(8):
        o.append_samples(dict(observation=obs, action=rng.integers(0, 6, (64, 256)), reward=rng.standard_normal((64, 256)).astype(np.float32),
                              done=rng.random((64, 256)) < 0.005))
    t0 = time.perf_counter()
    for _ in range(20):
        o.sample_batch(512)
    cpu_s = (time.perf_counter() - t0) / 20
    new = np.abs(rng.standard_normal(512)).astype(np.float32)
    cpu_su = 0.0                        # update timed on its own (sample_batch pairs with it but stays outside the clock;
    for _ in range(20):                 # round 1 subtracted two separately measured loops and printed a negative time)
        o.sample_batch(512)
        t0 = time.perf_counter()
        o.update_batch_priorities(new)
        cpu_su += time.perf_counter() - t0
    cpu_su /= 20
    out(kernel="oracle(CPU) replay.sample_batch(512)", ms=cpu_s * 1e3, update_ms=cpu_su * 1e3,
        speedup_sample=cpu_s / med, speedup_update=cpu_su / med_u)


def bench_conv(reps):
    import torch.nn.functional as F
    from rlpyt_b200.models import conv1_op
    from rlpyt_b200.models.conv2_op import conv2_relu
    g = torch.Generator(device="cuda").manual_seed(0)
    for N in (8192, 256):
        x = torch.relu(torch.randn(N, 16, 20, 20, device="cuda", generator=g))
        w = torch.randn(32, 16, 4, 4, device="cuda", generator=g) / 16
        b = torch.randn(32, device="cuda", generator=g)
        obs = torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
        w1 = torch.randn(16, 4, 8, 8, device="cuda", generator=g) / 16
        b1 = torch.randn(16, device="cuda", generator=g)
        t_tc2, _ = timeit(lambda: conv2_relu(x, w, b), reps, flush=False)
        t_cu2, _ = timeit(lambda: F.relu(F.conv2d(x, w, b, stride=2, padding=1)), reps, flush=False)
        conv1_op.FORWARD_IMPL = "tc"
        t_tc1, _ = timeit(lambda: conv1_op.conv1_u8_relu(w1, b1, obs, None), reps, flush=False)
        conv1_op.FORWARD_IMPL = "simt"
        t_si1, _ = timeit(lambda: conv1_op.conv1_u8_relu(w1, b1, obs, None), reps, flush=False)
        conv1_op.FORWARD_IMPL = "tc"
        t_cu1, _ = timeit(lambda: F.relu(F.conv2d(obs.float().mul_(1 / 255.), w1, b1, stride=4)), reps, flush=False)
        out(kernel="conv fwd", N=N, conv2_tc_us=t_tc2 * 1e6, conv2_cudnn_us=t_cu2 * 1e6, conv1_tc_us=t_tc1 * 1e6,
            conv1_simt_us=t_si1 * 1e6, conv1_cudnn_incl_convert_us=t_cu1 * 1e6)
        # backward pieces: weight gradient on tcgen05 vs the SIMT (layer 1) / cuDNN (layer 2) kernels
        from rlpyt_b200 import _lib
        from rlpyt_b200.models.conv2_op import wgrad_scratch
        o1 = torch.randn(N, 16, 20, 20, device="cuda", generator=g)
        g1 = torch.randn(N, 16, 20, 20, device="cuda", generator=g)
        o2 = torch.randn(N, 32, 10, 10, device="cuda", generator=g)
        g2 = (torch.randn(N, 32, 10, 10, device="cuda", generator=g) * (o2 > 0)).contiguous()
        gw1, gb1 = torch.empty_like(w1), torch.empty_like(b1)
        gw2, gb2 = torch.empty_like(w), torch.empty_like(b)
        sc = wgrad_scratch(x.device)
        sc1 = torch.empty(int(_lib.load().rl_conv1_u8_wgrad_scratch_bytes()) // 4, device="cuda")
        t_w1, _ = timeit(lambda: _lib.call("rl_conv1_u8_wgrad_tc", _lib.ptr(obs), None, _lib.ptr(o1), _lib.ptr(g1),
                                           _lib.ptr(gw1), _lib.ptr(gb1), N, 4, 84, 84, _lib.ptr(sc), _lib.stream()),
                         reps, flush=False)
        t_w1s, _ = timeit(lambda: _lib.call("rl_conv1_u8_wgrad", _lib.ptr(obs), None, _lib.ptr(o1), _lib.ptr(g1),
                                            _lib.ptr(gw1), _lib.ptr(gb1), N, 4, 84, 84, 1, _lib.ptr(sc1),
                                            _lib.stream()), reps, flush=False)
        t_w2, _ = timeit(lambda: _lib.call("rl_conv2_wgrad_tc", _lib.ptr(x), None, _lib.ptr(g2), _lib.ptr(gw2),
                                           _lib.ptr(gb2), N, 16, 20, 20, _lib.ptr(sc), _lib.stream()), reps,
                         flush=False)
        t_w2c, _ = timeit(lambda: torch.ops.aten.convolution_backward(
            g2, x, w, [32], [2, 2], [1, 1], [1, 1], False, [0, 0], 1, [False, True, True]), reps, flush=False)
        out(kernel="conv wgrad", N=N, conv1_tc_us=t_w1 * 1e6, conv1_simt_us=t_w1s * 1e6, conv2_tc_us=t_w2 * 1e6,
            conv2_cudnn_us=t_w2c * 1e6)


def bench_wgrad(reps):
    """tcgen05 weight-gradient kernels only (tuning runs: RLPYT_B200_WG_DEPTH=1..4)."""
    from rlpyt_b200 import _lib
    from rlpyt_b200.models.conv2_op import wgrad_scratch
    g = torch.Generator(device="cuda").manual_seed(0)
    N = 8192
    x = torch.relu(torch.randn(N, 16, 20, 20, device="cuda", generator=g))
    obs = torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    o1 = torch.randn(N, 16, 20, 20, device="cuda", generator=g)
    g1 = torch.randn(N, 16, 20, 20, device="cuda", generator=g)
    g2 = torch.randn(N, 32, 10, 10, device="cuda", generator=g)
    gw1, gb1 = torch.empty(16, 4, 8, 8, device="cuda"), torch.empty(16, device="cuda")
    gw2, gb2 = torch.empty(32, 16, 4, 4, device="cuda"), torch.empty(32, device="cuda")
    sc = wgrad_scratch(x.device)
    t1, _ = timeit(lambda: _lib.call("rl_conv1_u8_wgrad_tc", _lib.ptr(obs), None, _lib.ptr(o1), _lib.ptr(g1),
                                     _lib.ptr(gw1), _lib.ptr(gb1), N, 4, 84, 84, _lib.ptr(sc), _lib.stream()),
                   reps, flush=False)
    t2, _ = timeit(lambda: _lib.call("rl_conv2_wgrad_tc", _lib.ptr(x), None, _lib.ptr(g2), _lib.ptr(gw2),
                                     _lib.ptr(gb2), N, 16, 20, 20, _lib.ptr(sc), _lib.stream()), reps, flush=False)
    rows = torch.randperm(N, device="cuda", generator=g)
    t1r, _ = timeit(lambda: _lib.call("rl_conv1_u8_wgrad_tc", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(o1), _lib.ptr(g1),
                                      _lib.ptr(gw1), _lib.ptr(gb1), N, 4, 84, 84, _lib.ptr(sc), _lib.stream()),
                    reps, flush=False)
    out(kernel="conv wgrad tc", conv1_us=t1 * 1e6, conv1_rows_us=t1r * 1e6, conv2_us=t2 * 1e6)


BENCHES = {"wgrad": bench_wgrad, "returns": bench_returns, "gemm": bench_gemm, "replay": bench_replay, "conv": bench_conv}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=list(BENCHES))
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    out(device=torch.cuda.get_device_name(0), hbm_peak_gbs=HBM)
    for w in a.which:
        BENCHES[w](a.reps)
