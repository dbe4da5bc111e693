"""The sampler's step-loop protocol on the CPU: the REAL worker side (``sampling_process`` ->
``GpuResetCollector.collect_batch`` with forked processes, synthetic envs, the shared step buffer)
against a stand-in master that plays ``serve_actions`` without a GPU, for every handshake primitive the
sampler can be configured with (futex semaphores, spinning counters, spin-then-sleep).  Checks what
the master observes - rewards / dones / observations equal a host replay of the same seeded envs - and
that every semaphore is drained at the end of each batch (rlpyt/samplers/parallel/gpu/action_server.py:63-73)."""
import ctypes
import multiprocessing as mp

import numpy as np
import pytest

from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
from rlpyt_b200.samplers.buffer import StepBuffer
from rlpyt_b200.samplers.collections import TrajInfo
from rlpyt_b200.samplers.collectors import GpuResetCollector
from rlpyt_b200.samplers.parallel.gpu.sampler import sampling_process
from rlpyt_b200.utils.buffer import buffer_from_example
from rlpyt_b200.utils.collections import AttrDict
from rlpyt_b200.utils.seed import set_envs_seeds, set_seed
from rlpyt_b200.utils.synchronize import SpinSemaphore, SpinThenSleepSemaphore

ctx = mp.get_context("fork")
IMG, A = (4, 20, 20), 5
ENV_KW = dict(image_shape=IMG, n_actions=A, p_done=0.05, p_reward=0.3, pool_frames=8)


def _shared(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    return np.frombuffer(ctx.RawArray(ctypes.c_uint8, max(n, 1)), dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def _make_sem(kind):
    return {"futex": lambda: ctx.Semaphore(0), "spin": lambda: SpinSemaphore(ctx),
            "hybrid": lambda: SpinThenSleepSemaphore(ctx, spin_us=200.0)}[kind]()


@pytest.mark.parametrize("kind", ["futex", "spin", "hybrid"])
def test_worker_protocol_against_host_replay(kind):
    n_worker, n_envs, T, n_batches, seed = 3, 2, 12, 3, 7
    B = n_worker * n_envs
    step = StepBuffer(observation=_shared((B,) + IMG, np.uint8), action=_shared((B,), np.int64),
                      reward=_shared((B,), np.float32), done=_shared((B,), np.bool_))
    ctrl = AttrDict(quit=ctx.RawValue(ctypes.c_bool, False), barrier_in=ctx.Barrier(n_worker + 1),
                    barrier_out=ctx.Barrier(n_worker + 1), itr=ctx.RawValue(ctypes.c_long, 0))
    queue = ctx.Queue()
    probe_env = SyntheticAtariEnv(**ENV_KW)
    probe_env.reset()
    env_info_np = buffer_from_example(probe_env.step(0).env_info, (T, B), share_memory=True)   # as samplers/buffer.py
    obs_ready, act_ready = [_make_sem(kind) for _ in range(n_worker)], [_make_sem(kind) for _ in range(n_worker)]
    common = dict(EnvCls=SyntheticAtariEnv, env_kwargs=ENV_KW, batch_T=T, CollectorCls=GpuResetCollector,
                  TrajInfoCls=TrajInfo, traj_infos_queue=queue, ctrl=ctrl, max_decorrelation_steps=0, global_B=B)
    workers = []
    for w in range(n_worker):
        sl = slice(w * n_envs, (w + 1) * n_envs)
        wk = dict(rank=w, env_ranks=list(range(w * n_envs, (w + 1) * n_envs)), seed=seed + w, cpus=None, n_envs=n_envs,
                  step_buffer_np=step[sl], env_info_np=env_info_np[:, sl], sync=AttrDict(obs_ready=obs_ready[w], act_ready=act_ready[w]))
        workers.append(ctx.Process(target=sampling_process, kwargs=dict(common_kwargs=common, worker_kwargs=wk),
                                   daemon=True))
    for p in workers:
        p.start()
    # ---- host replay of the same seeded envs (what every worker does, in one process)
    replay_envs = []
    for w in range(n_worker):
        set_seed(seed + w)
        envs = [SyntheticAtariEnv(**ENV_KW) for _ in range(n_envs)]
        set_envs_seeds(envs, seed + w)
        replay_envs += envs
    expect_obs = np.stack([e.reset() for e in replay_envs])
    ctrl.barrier_out.wait()                               # workers have reset their envs
    assert np.array_equal(step.observation, expect_obs)
    rng = np.random.default_rng(0)
    try:
        for itr in range(n_batches):
            ctrl.itr.value = itr
            ctrl.barrier_in.wait()
            for t in range(T):                            # serve_actions without the GPU
                for s in obs_ready:
                    assert s.acquire(timeout=20)
                assert np.array_equal(step.observation, expect_obs)
                actions = rng.integers(0, A, B)
                step.action[:] = actions
                exp_r, exp_d = np.zeros(B, np.float32), np.zeros(B, bool)
                for b, env in enumerate(replay_envs):
                    o, r, d, info = env.step(actions[b])
                    if info.traj_done:
                        o = env.reset()
                    expect_obs[b], exp_r[b], exp_d[b] = o, r, d
                for s in act_ready:
                    s.release()
                last = (exp_r, exp_d)
            for s in obs_ready:
                assert s.acquire(timeout=20)
                assert not s.acquire(block=False)         # drained
            assert np.array_equal(step.observation, expect_obs)
            assert np.array_equal(step.reward, last[0]) and np.array_equal(step.done, last[1])
            for s in act_ready:
                assert not s.acquire(block=False)
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
    assert all(p.exitcode == 0 for p in workers)
