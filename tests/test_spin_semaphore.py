"""``SpinSemaphore`` (rlpyt_b200/utils/synchronize.py): semaphore semantics across forked processes -
the step loop's obs_ready / act_ready protocol (rlpyt/samplers/parallel/gpu/action_server.py:37-62:
T+1 releases of obs_ready, T of act_ready per batch, drained at the end) with data handed over through
shared memory."""
import multiprocessing as mp
import time

import numpy as np
import pytest

from rlpyt_b200.utils.synchronize import SpinSemaphore


def test_counting_semantics_single_process():
    s = SpinSemaphore()
    assert not s.acquire(block=False)
    s.release()
    s.release()
    assert s.acquire() and s.acquire(block=False) and not s.acquire(block=False)
    assert not s.acquire(timeout=0.05)


def _worker(obs_ready, act_ready, data, T, out):
    obs_ready.release()                                  # observation(0) is in place
    total = 0
    for t in range(T):
        act_ready.acquire()
        total += int(data[0])                            # the "action" the master wrote before releasing
        data[1] = t + 1                                  # the "observation" for the next step
        obs_ready.release()
    out.put(total)


@pytest.mark.parametrize("n_workers", [1, 3])
def test_step_protocol_across_processes(n_workers):
    ctx = mp.get_context("fork")
    T = 300
    obs = [SpinSemaphore(ctx) for _ in range(n_workers)]
    act = [SpinSemaphore(ctx) for _ in range(n_workers)]
    data = [np.frombuffer(ctx.RawArray("q", 2), dtype=np.int64) for _ in range(n_workers)]
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(obs[i], act[i], data[i], T, out)) for i in range(n_workers)]
    for p in procs:
        p.start()
    t0 = time.perf_counter()
    for t in range(T):
        for i in range(n_workers):
            assert obs[i].acquire(timeout=30)
            assert data[i][1] == t                       # the worker's write is visible once acquired
            data[i][0] = t
        for s in act:
            s.release()
    for s in obs:
        assert s.acquire(timeout=30)
        assert not s.acquire(block=False)                # drained, as the reference asserts
    dt = time.perf_counter() - t0
    totals = sorted(out.get(timeout=30) for _ in procs)
    for p in procs:
        p.join(timeout=10)
        assert p.exitcode == 0
    assert totals == [T * (T - 1) // 2] * n_workers
    for s in act:
        assert not s.acquire(block=False)
    assert dt < 60
