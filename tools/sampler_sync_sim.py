#!/usr/bin/env python
"""CPU-only simulation of the sampler's step loop: the REAL forked worker processes (synthetic Atari envs,
GpuResetCollector, shared step buffer) against the REAL master loops (GpuSampler.serve_actions and
AlternatingSampler.serve_actions) with a stand-in step engine that busy-waits for the measured GPU phase
of a step (H2D + graph replay + D2H sync, ~340 us full batch) instead of running it.  Prints ms per
128-step batch for every (master loop, handshake primitive) pair.  Numbers depend on the host (cores, SMT,
scheduler); they rank the options and size the handshake overhead, they are not bench values.

    python tools/sampler_sync_sim.py [--workers 6] [--envs 18] [--T 128] [--gpu-us 340] [--half-gpu-us 220]
"""
import argparse
import ctypes
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlpyt_b200.envs.synthetic import SyntheticAtariEnv  # noqa: E402
from rlpyt_b200.samplers.buffer import StepBuffer  # noqa: E402
from rlpyt_b200.samplers.collections import BatchSpec, TrajInfo  # noqa: E402
from rlpyt_b200.samplers.collectors import GpuResetCollector  # noqa: E402
from rlpyt_b200.samplers.parallel.gpu.alternating_sampler import AlternatingSampler  # noqa: E402
from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler, sampling_process  # noqa: E402
from rlpyt_b200.utils.buffer import buffer_from_example  # noqa: E402
from rlpyt_b200.utils.collections import AttrDict  # noqa: E402
from rlpyt_b200.utils.synchronize import SpinSemaphore, SpinThenSleepSemaphore  # noqa: E402

ctx = mp.get_context("fork")
IMG, A = (4, 84, 84), 6
ENV_KW = dict(image_shape=IMG, n_actions=A, p_done=1 / 500., p_reward=0.04)


def shared(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer(ctx.RawArray(ctypes.c_uint8, max(n, 1)), dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class FakeRollout:
    def __init__(self, step_np, gpu_us):
        self.step_np, self.gpu_s = step_np, gpu_us * 1e-6
        self.rng = np.random.default_rng(0)

    side_stream = None

    def step(self, t, zero_inputs_on_done, blank_done_rows=False, obs_done=False):
        end = time.perf_counter() + self.gpu_s
        while time.perf_counter() < end:      # the master thread is busy in the driver for this long
            pass
        self.step_np.action[:] = self.rng.integers(0, A, len(self.step_np.action))

    # the alternating master's surface: upload and act issued separately, completion polled or waited for.  The device
    # phase is modelled as a deadline: H2D 45 % of the half step, agent.step the rest, strictly one after the other.
    def upload_worker_rows(self, k, i):
        pass

    def upload_async(self, k, zero_inputs_on_done, obs_done=False):
        self._ready_at = max(getattr(self, "_ready_at", 0.0), time.perf_counter()) + 0.45 * self.gpu_s

    def act_async(self, k, blank_done_rows=False):
        self._ready_at = max(getattr(self, "_ready_at", 0.0), time.perf_counter()) + 0.55 * self.gpu_s

    def act_done(self):
        return time.perf_counter() >= getattr(self, "_ready_at", 0.0)

    def wait(self):
        while time.perf_counter() < getattr(self, "_ready_at", 0.0):
            pass
        self.step_np.action[:] = self.rng.integers(0, A, len(self.step_np.action))

    def finish(self):
        pass

    def zero_inputs_where_done(self):
        pass

    def end_batch(self):
        pass


def run(kind, alternating, args):
    n_worker, n_envs, T = args.workers, args.envs, args.T
    B = n_worker * n_envs
    step = StepBuffer(observation=shared((B,) + IMG, np.uint8), action=shared((B,), np.int64),
                      reward=shared((B,), np.float32), done=shared((B,), np.bool_))
    ctrl = AttrDict(quit=ctx.RawValue(ctypes.c_bool, False), barrier_in=ctx.Barrier(n_worker + 1),
                    barrier_out=ctx.Barrier(n_worker + 1), itr=ctx.RawValue(ctypes.c_long, 0))
    make = {"futex": lambda: ctx.Semaphore(0), "spin": lambda: SpinSemaphore(ctx),
            "hybrid": lambda: SpinThenSleepSemaphore(ctx)}[kind]
    obs_ready, act_ready = [make() for _ in range(n_worker)], [make() for _ in range(n_worker)]
    env = SyntheticAtariEnv(**ENV_KW)
    env.reset()
    env_info_np = buffer_from_example(env.step(0).env_info, (T, B), share_memory=True)
    common = dict(EnvCls=SyntheticAtariEnv, env_kwargs=ENV_KW, batch_T=T, CollectorCls=GpuResetCollector,
                  TrajInfoCls=TrajInfo, traj_infos_queue=ctx.Queue(), ctrl=ctrl, max_decorrelation_steps=0, global_B=B)
    workers = []
    for w in range(n_worker):
        sl = slice(w * n_envs, (w + 1) * n_envs)
        wk = dict(rank=w, env_ranks=list(range(sl.start, sl.stop)), seed=w, cpus=None, n_envs=n_envs,
                  step_buffer_np=step[sl], env_info_np=env_info_np[:, sl],
                  sync=AttrDict(obs_ready=obs_ready[w], act_ready=act_ready[w]))
        workers.append(ctx.Process(target=sampling_process, kwargs=dict(common_kwargs=common, worker_kwargs=wk),
                                   daemon=True))
    for p in workers:
        p.start()
    Cls = AlternatingSampler if alternating else GpuSampler
    s = Cls(EnvCls=SyntheticAtariEnv, env_kwargs=ENV_KW, batch_T=T, batch_B=B)
    s.batch_spec, s.mid_batch_reset, s.device = BatchSpec(T, B), True, torch.device("cpu")
    s.agent = type("Agent", (), {"reset_one": lambda self, idx: None})()
    s.sync = AttrDict(obs_ready=obs_ready, act_ready=act_ready)
    if alternating:
        hw, hb = n_worker // 2, B // 2
        s.halves = (slice(0, hb), slice(hb, B))
        s.obs_ready_pair, s.act_ready_pair = (obs_ready[:hw], obs_ready[hw:]), (act_ready[:hw], act_ready[hw:])
        s.rollouts = [FakeRollout(step[sl], args.half_gpu_us) for sl in s.halves]
    else:
        s.host = dict(step_np=step)
        s.rollout = FakeRollout(step, args.gpu_us)
    ctrl.barrier_out.wait()
    times = []
    try:
        for itr in range(args.batches + 1):
            ctrl.barrier_in.wait()
            t0 = time.perf_counter()
            if alternating:
                s.serve_actions(itr)
            else:
                serve_standard(s, T)
            times.append(time.perf_counter() - t0)
            ctrl.barrier_out.wait()
    finally:
        ctrl.quit.value = True
        try:
            ctrl.barrier_in.wait(timeout=10)
        except Exception:
            pass
        for p in workers:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
    return float(np.median(times[1:])) * 1e3


def serve_standard(s, T):
    """GpuSampler.serve_actions without its CUDA stream synchronisation (same handshake order)."""
    obs_ready, act_ready, ro = s.sync.obs_ready, s.sync.act_ready, s.rollout
    for t in range(T):
        for sem in obs_ready:
            sem.acquire()
        ro.step(t, zero_inputs_on_done=True)
        for sem in act_ready:
            sem.release()
    for sem in obs_ready:
        sem.acquire()
    ro.finish()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=6)
    ap.add_argument("--envs", type=int, default=18)
    ap.add_argument("--T", type=int, default=128)
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--gpu-us", type=float, default=340.0)
    ap.add_argument("--half-gpu-us", type=float, default=220.0)
    args = ap.parse_args()
    print(f"host: {os.cpu_count()} logical cpus; {args.workers} workers x {args.envs} envs, T={args.T}, "
          f"GPU phase {args.gpu_us} us (full) / {args.half_gpu_us} us (half)")
    for alternating in (False, True):
        for kind in ("futex", "hybrid", "spin"):
            ms = run(kind, alternating, args)
            print(f"{'alternating' if alternating else 'standard   '} {kind:6s}: {ms:7.1f} ms per batch "
                  f"({ms / args.T * 1e3:6.0f} us per step)")
