"""The worker's batch loop (rlpyt_b200/samplers/collectors.py ``GpuResetCollector.collect_batch``: reward / done / env_info
written once per time step, cached observation rows) against the reference loop written out per environment
(rlpyt/samplers/parallel/gpu/collectors.py:25-48): after every step the step buffer holds the same observation, reward
and done, the env_info log rows and the completed trajectory infos are the same."""
import threading

import numpy as np

from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
from rlpyt_b200.samplers.buffer import StepBuffer
from rlpyt_b200.samplers.collections import TrajInfo
from rlpyt_b200.samplers.collectors import GpuResetCollector
from rlpyt_b200.utils.buffer import buffer_from_example
from rlpyt_b200.utils.collections import AttrDict

ENV_KW = dict(image_shape=(4, 12, 12), n_actions=5, p_done=0.15, p_reward=0.5)


def _envs(n, seed):
    envs = [SyntheticAtariEnv(**ENV_KW) for _ in range(n)]
    for i, e in enumerate(envs):
        e.seed(seed + i)
    return envs


def test_collect_batch_matches_the_per_environment_reference_loop():
    n, T, seed = 5, 12, 7
    rng = np.random.default_rng(0)
    actions = rng.integers(0, 5, size=(2 * T, n))
    # ---- reference loop (gpu/collectors.py:25-48), two batches
    envs = _envs(n, seed)
    obs = np.stack([e.reset() for e in envs])
    want_steps, want_infos, want_trajs = [], [], []
    traj = [TrajInfo() for _ in range(n)]
    for t in range(2 * T):
        rew, done, infos = np.zeros(n, np.float32), np.zeros(n, bool), []
        for b, env in enumerate(envs):
            o, r, d, info = env.step(actions[t, b])
            traj[b].step(obs[b], actions[t, b], r, d, None, info)
            if getattr(info, "traj_done", d):
                want_trajs.append(dict(traj[b].terminate(o)))
                traj[b] = TrajInfo()
                o = env.reset()
            obs[b] = o
            rew[b], done[b] = r, d
            infos.append(info)
        want_steps.append((obs.copy(), rew, done))
        want_infos.append(infos)
    # ---- the collector, driven step by step by a stand-in master
    envs = _envs(n, seed)
    step = StepBuffer(observation=np.stack([e.reset() for e in envs]), action=np.zeros(n, np.int64),
                      reward=np.zeros(n, np.float32), done=np.zeros(n, bool))
    probe = envs[0].__class__(**ENV_KW)
    probe.reset()
    env_info_np = buffer_from_example(probe.step(0).env_info, (T, n))
    sync = AttrDict(obs_ready=threading.Semaphore(0), act_ready=threading.Semaphore(0))
    col = GpuResetCollector(rank=0, envs=envs, env_info_np=env_info_np, batch_T=T, TrajInfoCls=TrajInfo, sync=sync,
                            step_buffer_np=step)
    got_trajs, got_steps, got_infos = [], [], []
    traj_infos = [TrajInfo() for _ in range(n)]
    for batch in range(2):
        out = {}

        def work():
            out["r"] = col.collect_batch(traj_infos, batch)
        th = threading.Thread(target=work)
        th.start()
        sync.obs_ready.acquire()                              # the observation of the batch's first step
        for t in range(T):
            step.action[:] = actions[batch * T + t]
            sync.act_ready.release()
            sync.obs_ready.acquire()
            got_steps.append((step.observation.copy(), step.reward.copy(), step.done.copy()))
        th.join()
        traj_infos, completed = out["r"]
        got_trajs += [dict(c) for c in completed]
        got_infos.append({f: getattr(env_info_np, f).copy() for f in env_info_np._fields})
    for t, ((o, r, d), (o2, r2, d2)) in enumerate(zip(want_steps, got_steps)):
        assert np.array_equal(o, o2) and np.array_equal(r, r2) and np.array_equal(d, d2), t
    for batch in range(2):
        for f in env_info_np._fields:
            want = np.array([[getattr(i, f) for i in want_infos[batch * T + t]] for t in range(T)])
            assert np.array_equal(got_infos[batch][f], want), (batch, f)
    assert len(got_trajs) == len(want_trajs) > 0
    for a, b in zip(got_trajs, want_trajs):
        assert a.keys() == b.keys()
        for k in a:
            assert a[k] == b[k] and type(a[k]) is type(b[k]), k      # same values AND the reference loop's scalar types
