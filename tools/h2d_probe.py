#!/usr/bin/env python
"""H2D bandwidth of the step-buffer candidates: torch pinned, fork-shared + cudaHostRegister
(touched by the master / by a child on another core), pageable."""
import ctypes, multiprocessing as mp, os, time
import numpy as np, torch
N = 256 * 4 * 84 * 84
dev = torch.empty(N, dtype=torch.uint8, device="cuda")

def bw(t, label, reps=20):
    for _ in range(3): dev.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): dev.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label:50s} {N/dt/1e9:7.2f} GB/s  {dt*1e6:8.1f} us", flush=True)

os.system("nvidia-smi topo -m | head -8; numactl -H 2>/dev/null | head -6; lscpu | grep -i numa")
bw(torch.empty(N, dtype=torch.uint8).pin_memory(), "torch pin_memory")
bw(torch.empty(N, dtype=torch.uint8), "pageable")
ctx = mp.get_context("fork")
def shared(touch_in_child, cpu=None):
    raw = ctx.RawArray(ctypes.c_char, N)
    arr = np.frombuffer(raw, dtype=np.uint8)
    if touch_in_child:
        def f():
            if cpu is not None: os.sched_setaffinity(0, [cpu])
            arr[:] = 7
        p = ctx.Process(target=f); p.start(); p.join()
    else:
        arr[:] = 7
    rc = torch.cuda.cudart().cudaHostRegister(arr.ctypes.data, arr.nbytes, 0)
    return torch.from_numpy(arr), rc
t, rc = shared(False); bw(t, f"RawArray touched by master + HostRegister rc={int(rc)}")
for cpu in (0, 32, 64, 96, 127):
    try:
        t, rc = shared(True, cpu); bw(t, f"RawArray touched by child on cpu {cpu} + HostRegister rc={int(rc)}")
    except Exception as e:
        print("cpu", cpu, e)
print("master affinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
