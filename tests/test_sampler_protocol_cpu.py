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


class _FakeRollout:
    """Stands in for DeviceRollout on the CPU: 'agent.step' = seeded random actions written to the step buffer;
    records the order of events so the alternating master loop can be checked.  ``upload_worker_rows`` snapshots a
    worker's rows of the step buffer at the moment the master issues their DMA (it does so per worker, as soon as that
    worker has signalled - possibly while the OTHER half's agent.step is in flight): the rows must already be the
    observations of step k."""

    def __init__(self, step_np, rng, log, tag, chunks):
        self.step_np, self.rng, self.log, self.tag = step_np, rng, log, tag
        self.chunks = chunks                              # (start, n) of each worker within this half
        self.obs_seen = []
        self._cur = {}
        self._polls = 0

    side_stream = None

    def upload_worker_rows(self, k, i):
        start, n = self.chunks[i]
        buf = self._cur.setdefault(k, np.zeros_like(self.step_np.observation))
        buf[start:start + n] = self.step_np.observation[start:start + n]
        self.log.append((self.tag, "rows", k, i))

    def upload_async(self, t, zero_inputs_on_done, obs_done=False):
        if obs_done:                                      # RLPYT_B200_SAMPLER_CHUNKED=1: every worker's rows were uploaded one by one
            assert sorted(e[3] for e in self.log if e[:3] == (self.tag, "rows", t)) == list(range(len(self.chunks)))
            self.obs_seen.append(self._cur.pop(t))
        else:                                             # one DMA of the half's whole step buffer, now
            assert not any(e[:3] == (self.tag, "rows", t) for e in self.log)
            self.obs_seen.append(self.step_np.observation.copy())
        self.log.append((self.tag, "up", t))

    def act_async(self, t, blank_done_rows=False):
        self.step_np.action[:] = self.rng.integers(0, A, len(self.step_np.action))
        self.log.append((self.tag, t))
        self._polls = 0

    def act_done(self):
        self._polls += 1                                  # the 'device' takes a few polls: the master keeps serving the other half
        return self._polls > 3

    def wait(self):
        pass

    def finish(self):
        self.obs_seen.append(self.step_np.observation.copy())
        self.log.append((self.tag, "finish"))

    def zero_inputs_where_done(self):
        pass

    def end_batch(self):
        pass


@pytest.mark.parametrize("kind,chunked,poll", [("futex", "0", "once"), ("hybrid", "0", "once"), ("futex", "1", "spin"),
                                               ("hybrid", "1", "yield"), ("futex", "1", "once")])
def test_alternating_master_loop_against_real_workers(kind, chunked, poll, monkeypatch):
    """``AlternatingSampler.serve_actions`` (the only new logic of that class) with the real forked worker loop:
    strict (half 0, half 1) alternation per step, no deadlock, all handshakes drained, and each half's
    observations equal a host replay of its envs under the actions the master chose."""
    import torch
    from rlpyt_b200.samplers.collections import BatchSpec
    from rlpyt_b200.samplers.parallel.gpu.alternating_sampler import AlternatingSampler
    monkeypatch.setenv("RLPYT_B200_SAMPLER_CHUNKED", chunked)     # per-worker uploads or one upload per half
    monkeypatch.setenv("RLPYT_B200_SAMPLER_POLL", poll)           # how the master looks at the stepping half
    n_worker, n_envs, T, seed = 4, 2, 10, 11
    B = n_worker * n_envs
    step = StepBuffer(observation=_shared((B,) + IMG, np.uint8), action=_shared((B,), np.int64),
                      reward=_shared((B,), np.float32), done=_shared((B,), np.bool_))
    ctrl = AttrDict(quit=ctx.RawValue(ctypes.c_bool, False), barrier_in=ctx.Barrier(n_worker + 1),
                    barrier_out=ctx.Barrier(n_worker + 1), itr=ctx.RawValue(ctypes.c_long, 0))
    queue = ctx.Queue()
    probe_env = SyntheticAtariEnv(**ENV_KW)
    probe_env.reset()
    env_info_np = buffer_from_example(probe_env.step(0).env_info, (T, B), share_memory=True)
    obs_ready, act_ready = [_make_sem(kind) for _ in range(n_worker)], [_make_sem(kind) for _ in range(n_worker)]
    common = dict(EnvCls=SyntheticAtariEnv, env_kwargs=ENV_KW, batch_T=T, CollectorCls=GpuResetCollector,
                  TrajInfoCls=TrajInfo, traj_infos_queue=queue, ctrl=ctrl, max_decorrelation_steps=0, global_B=B)
    workers = []
    for w in range(n_worker):
        sl = slice(w * n_envs, (w + 1) * n_envs)
        wk = dict(rank=w, env_ranks=list(range(sl.start, sl.stop)), seed=seed + w, cpus=None, n_envs=n_envs,
                  step_buffer_np=step[sl], env_info_np=env_info_np[:, sl],
                  sync=AttrDict(obs_ready=obs_ready[w], act_ready=act_ready[w]))
        workers.append(ctx.Process(target=sampling_process, kwargs=dict(common_kwargs=common, worker_kwargs=wk),
                                   daemon=True))
    for p in workers:
        p.start()
    # the sampler object, wired by hand where initialize() would need CUDA
    s = AlternatingSampler(EnvCls=SyntheticAtariEnv, env_kwargs=ENV_KW, batch_T=T, batch_B=B)
    s.batch_spec, s.mid_batch_reset, s.device = BatchSpec(T, B), True, torch.device("cpu")
    s.agent = type("Agent", (), {"reset_one": lambda self, idx: None})()
    s.sync = AttrDict(obs_ready=obs_ready, act_ready=act_ready)
    half_w, half_B = n_worker // 2, B // 2
    s.halves = (slice(0, half_B), slice(half_B, B))
    s.obs_ready_pair = (obs_ready[:half_w], obs_ready[half_w:])
    s.act_ready_pair = (act_ready[:half_w], act_ready[half_w:])
    log = []
    s.rollouts = [_FakeRollout(step[sl], np.random.default_rng(i), log, i, [(w * n_envs, n_envs) for w in range(half_w)])
                  for i, sl in enumerate(s.halves)]
    # host replay
    replay_envs = []
    for w in range(n_worker):
        set_seed(seed + w)
        envs = [SyntheticAtariEnv(**ENV_KW) for _ in range(n_envs)]
        set_envs_seeds(envs, seed + w)
        replay_envs += envs
    expect = np.stack([e.reset() for e in replay_envs])
    ctrl.barrier_out.wait()
    try:
        ctrl.barrier_in.wait()
        s.serve_actions(0)
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
    want = [(h, t) for t in range(T) for h in (0, 1)] + [(0, "finish"), (1, "finish")]
    assert [e for e in log if len(e) == 2] == want            # strict (half 0, half 1) alternation of the agent steps
    ups = [e for e in log if len(e) == 3]
    assert sorted(ups) == sorted((h, "up", t) for t in range(T) for h in (0, 1))
    for h in (0, 1):                                          # every upload precedes its own agent step
        for t in range(T):
            assert log.index((h, "up", t)) < log.index((h, t))
    # replay: actions of each half come from the same seeded generators, in the same order
    rngs = [np.random.default_rng(0), np.random.default_rng(1)]
    for t in range(T + 1):
        for h, sl in enumerate(s.halves):
            assert np.array_equal(s.rollouts[h].obs_seen[t], expect[sl])
            if t < T:
                acts = rngs[h].integers(0, A, half_B)
                for i, b in enumerate(range(sl.start, sl.stop)):
                    o, r, d, info = replay_envs[b].step(acts[i])
                    if info.traj_done:
                        o = replay_envs[b].reset()
                    expect[b] = o
