#!/usr/bin/env python
"""What one HALF step of the alternating sampler costs on the device side, piece by piece (B/2 = 128 envs of (4,84,84)):
upload graph (H2D of 3.6 MB + the small fields), act graph (agent.step + D2H of the actions), each alone and the two
overlapped on two streams the way serve_actions issues them; wall clock of the host calls and CUDA-event times.
Never a bench value."""
import json
import os
import sys
import time
from collections import namedtuple

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlpyt_b200.agents.pg.atari import AtariFfAgent  # noqa: E402
from rlpyt_b200.samplers.buffer import build_samples_buffer  # noqa: E402
from rlpyt_b200.samplers.collections import BatchSpec  # noqa: E402
from rlpyt_b200.samplers.rollout import DeviceRollout  # noqa: E402
from rlpyt_b200.envs.synthetic import SyntheticAtariEnv  # noqa: E402

dev = torch.device("cuda", 0)
env = SyntheticAtariEnv()
agent = AtariFfAgent()
agent.initialize(env.spaces)
agent.to_device(0)
agent.sample_mode(0)
T, B = 8, 256
samples, host, _ = build_samples_buffer(agent, env, BatchSpec(T, B), True, device=dev, share_host=False)
halves = (slice(0, B // 2), slice(B // 2, B))
ros = []
for sl in halves:
    h = dict(step_np=host["step_np"][sl], step_pyt=host["step_pyt"][sl], all_action=host["all_action"][:, sl],
             all_reward=host["all_reward"][:, sl], pinned=True)
    ros.append(DeviceRollout(samples[:, sl], h, agent, dev, stream=torch.cuda.Stream(dev)))
for ro in ros:      # eager batch, then capture
    for t in range(T):
        ro.upload_async(t, True); ro.act_async(t); ro.wait()
    ro.end_batch()
for ro in ros:
    for t in range(T):
        ro.upload_async(t, True); ro.act_async(t); ro.wait()
out = {}


def wall(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


a, b = ros
out["upload_then_wait_us"] = wall(lambda: (a.upload_async(1, True), a.wait()))
out["act_then_wait_us"] = wall(lambda: (a.act_async(1), a.wait()))
out["upload_act_wait_us"] = wall(lambda: (a.upload_async(1, True), a.act_async(1), a.wait()))
# the serve_actions pattern: act(this) ; upload(other) ; wait(this)
def pattern():
    a.act_async(1); b.upload_async(1, True); a.wait()
    b.act_async(1); a.upload_async(2, True); b.wait()
out["two_half_steps_pattern_us"] = wall(pattern)
# host cost of the launches alone (no wait in between)
def launches():
    a.act_async(1); b.upload_async(1, True)
t0 = time.perf_counter()
for _ in range(200):
    launches()
out["host_issue_act_plus_upload_us"] = (time.perf_counter() - t0) / 200 * 1e6
torch.cuda.synchronize()
# raw copies
obs_h = a.step_pyt.observation
obs_d = samples.env.observation[0][halves[0]]
out["raw_h2d_3p6MB_us"] = wall(lambda: (obs_d.copy_(obs_h, non_blocking=True), torch.cuda.current_stream().synchronize()))
act_h, act_d = a.step_pyt.action, samples.agent.action[0][halves[0]]
out["raw_d2h_actions_us"] = wall(lambda: (act_h.copy_(act_d, non_blocking=True), torch.cuda.current_stream().synchronize()))
out["empty_sync_us"] = wall(lambda: torch.cuda.current_stream().synchronize())
print(json.dumps(out))
