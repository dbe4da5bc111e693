#!/usr/bin/env python
"""Fine-grained timing of the sampler's per-step device work inside a live GpuSampler loop."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rlpyt_b200.agents.pg.atari import AtariFfAgent
from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler

workers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.backends.cudnn.benchmark = True
sampler = GpuSampler(EnvCls=SyntheticAtariEnv, env_kwargs=bench.ENV_KW, batch_T=128, batch_B=256, max_decorrelation_steps=0)
agent = AtariFfAgent()
from rlpyt_b200.utils.affinity import make_affinity
aff = make_affinity(0, workers) if os.environ.get("PIN", "1") == "1" else dict(cuda_idx=0, workers_cpus=[None] * workers, set_affinity=False)
print("affinity master", aff.get("master_cpus"), "workers", aff["workers_cpus"][:4], "...")
sampler.initialize(agent, affinity=aff, seed=1, bootstrap_value=True)
agent.to_device(0)
ro = sampler.rollout
for i in range(2):
    sampler.obtain_samples(i)
acc = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = fn(*a, **k); torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    return w
s = ro.samples
orig_upload = ro.upload
def upload(k, zero_inputs_on_done, obs_done=False):
    obs_dst = s.env.observation[k] if k < ro.T else ro.obs_extra
    t0 = time.perf_counter()
    if not obs_done: obs_dst.copy_(ro.step_pyt.observation, non_blocking=True)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    acc["obs_copy_issue"] = acc.get("obs_copy_issue", 0) + t1 - t0; acc["obs_copy_wait"] = acc.get("obs_copy_wait", 0) + t2 - t1
    t0 = time.perf_counter()
    ro.all_reward[k].copy_(ro.step_pyt.reward, non_blocking=True)
    ro.done_step.copy_(ro.step_pyt.done, non_blocking=True)
    if k >= 1: s.env.done[k - 1].copy_(ro.done_step, non_blocking=True)
    ro.in_reward.copy_(ro.all_reward[k], non_blocking=True)
    if zero_inputs_on_done:
        ro.in_action.masked_fill_(ro.done_step, 0); ro.in_reward.masked_fill_(ro.done_step, 0)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    acc["small_issue"] = acc.get("small_issue", 0) + t1 - t0; acc["small_wait"] = acc.get("small_wait", 0) + t2 - t1
    return obs_dst
ro.upload = upload
ro.act = timed("act", ro.act)
t0 = time.perf_counter(); sampler.obtain_samples(3); tot = time.perf_counter() - t0
print(f"workers={workers} obtain_samples {tot*1e3:.1f} ms; per-step us:", {k: round(v / 129 * 1e6, 1) for k, v in acc.items()},
      "rest(us/step)", round((tot - sum(acc.values())) / 128 * 1e6, 1), flush=True)
# copy speed of the same buffer when nobody touched it recently
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): s.env.observation[0].copy_(ro.step_pyt.observation, non_blocking=True)
torch.cuda.synchronize(); print("idle obs copy us", (time.perf_counter() - t0) / 20 * 1e6)
# after the master itself dirties it
ro.step_np.observation[:] = 3
torch.cuda.synchronize(); t0 = time.perf_counter(); s.env.observation[0].copy_(ro.step_pyt.observation, non_blocking=True); torch.cuda.synchronize()
print("obs copy right after master wrote it us", (time.perf_counter() - t0) * 1e6)
sampler.shutdown()
