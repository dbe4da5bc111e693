#!/usr/bin/env python
"""Where the device half of a sampler step goes (B = 256 and 128, (4,84,84) frames): H2D of the step buffer from pinned
memory, agent.step as a CUDA graph, D2H of the actions + stream sync - CUDA events / wall clock.  Never a bench value."""
import json
import sys
import time
from collections import namedtuple

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from rlpyt_b200.agents.pg.atari import AtariFfAgent  # noqa: E402

Spaces = namedtuple("Spaces", "observation action")
agent = AtariFfAgent()
agent.initialize(Spaces(namedtuple("O", "shape")((4, 84, 84)), namedtuple("Ac", "n")(6)))
agent.to_device(0)
agent.sample_mode(0)
out = {}
for B in (256, 128):
    host = torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8).pin_memory()
    host_act = torch.zeros(B, dtype=torch.int64).pin_memory()
    obs = torch.empty((B, 4, 84, 84), dtype=torch.uint8, device="cuda")
    pa, pr = torch.zeros(B, dtype=torch.int64, device="cuda"), torch.zeros(B, device="cuda")
    for _ in range(3):
        step = agent.step(obs, pa, pr)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step = agent.step(obs, pa, pr)

    def ev(fn, reps=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    r = {"h2d_us": ev(lambda: obs.copy_(host, non_blocking=True)), "agent_step_graph_us": ev(g.replay),
         "agent_step_eager_us": ev(lambda: agent.step(obs, pa, pr), 20)}
    t0 = time.perf_counter()
    for _ in range(200):
        obs.copy_(host, non_blocking=True)
        g.replay()
        host_act.copy_(step.action, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    r["h2d_graph_d2h_sync_wall_us"] = (time.perf_counter() - t0) / 200 * 1e6
    out[f"B={B}"] = r
print(json.dumps(out))
