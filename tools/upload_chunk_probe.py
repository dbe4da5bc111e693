#!/usr/bin/env python
"""Host-side cost and device-side completion of the per-worker observation uploads: 7 x 0.5 MB out of a fork-shared,
cudaHostRegister'ed step buffer (exactly what the samplers use) through rl_upload_async vs Tensor.copy_ vs ONE 3.6 MB copy.
Never a bench value."""
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlpyt_b200 import _lib  # noqa: E402
from rlpyt_b200.samplers.buffer import pin_shared  # noqa: E402

ctx = mp.get_context("fork")
B, ROW = 128, 4 * 84 * 84
raw = ctx.RawArray(ctypes.c_char, B * ROW)
host = np.frombuffer(raw, dtype=np.uint8).reshape(B, ROW)
host[:] = 3
torch.zeros(1, device="cuda")
ok = pin_shared(host)
dev = torch.empty((B, ROW), dtype=torch.uint8, device="cuda")
host_t = torch.from_numpy(host)
pinned_t = torch.empty((B, ROW), dtype=torch.uint8).pin_memory()
fn = _lib.load().rl_upload_async
stream = torch.cuda.current_stream().cuda_stream
chunks = [(i * 18, 18 if i < 6 else B - 108) for i in range(7)]
out = {"pin_shared_ok": bool(ok)}


def bench(issue, reps=200):
    for _ in range(10):
        issue()
    torch.cuda.synchronize()
    t_issue = t_total = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        issue()
        t1 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        t2 = time.perf_counter()
        t_issue += t1 - t0
        t_total += t2 - t0
    return {"host_issue_us": t_issue / reps * 1e6, "until_done_us": t_total / reps * 1e6}


def abi_chunks(src_np):
    base_s, base_d = int(src_np.ctypes.data), int(dev.data_ptr())
    def f():
        for s, n in chunks:
            fn(base_d + s * ROW, base_s + s * ROW, n * ROW, stream)
    return f


def torch_chunks(src_t):
    def f():
        for s, n in chunks:
            dev[s:s + n].copy_(src_t[s:s + n], non_blocking=True)
    return f


out["abi_7_chunks_registered"] = bench(abi_chunks(host))
out["abi_7_chunks_torch_pinned"] = bench(abi_chunks(pinned_t.numpy()))
out["torch_7_chunks_registered"] = bench(torch_chunks(host_t))
out["torch_7_chunks_torch_pinned"] = bench(torch_chunks(pinned_t))
out["torch_one_copy_registered"] = bench(lambda: dev.copy_(host_t, non_blocking=True))
out["abi_one_copy_registered"] = bench(lambda: fn(int(dev.data_ptr()), int(host.ctypes.data), B * ROW, stream))
out["torch_one_copy_torch_pinned"] = bench(lambda: dev.copy_(pinned_t, non_blocking=True))
print(json.dumps(out))
