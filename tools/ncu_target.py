#!/usr/bin/env python
"""Small, deterministic targets for ncu (never a bench value):
    ncu ... python tools/ncu_target.py ppo      # one PPO optimize_agent iteration on resident samples
    ncu ... python tools/ncu_target.py gae      # GAE streaming kernel at [128, 2^20] + tscan at [128,256]
    ncu ... python tools/ncu_target.py replay   # sum-tree sample/update + frame gather
Profiling is limited to the region between cudaProfilerStart/Stop (use --profile-from-start off).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def resident_samples(T, B, A, dev="cuda"):
    from rlpyt_b200.samplers.collections import Samples, AgentSamplesBsv, EnvSamples
    from rlpyt_b200.agents.pg.base import AgentInfo
    from rlpyt_b200.distributions.categorical import DistInfo
    g = torch.Generator(device=dev).manual_seed(0)
    obs = torch.randint(0, 256, (T, B) + bench.IMAGE, dtype=torch.uint8, device=dev, generator=g)
    all_action = torch.randint(0, A, (T + 1, B), device=dev, generator=g)
    all_reward = torch.randn(T + 1, B, device=dev, generator=g)
    prob = torch.softmax(torch.randn(T, B, A, device=dev, generator=g), -1)
    return Samples(
        agent=AgentSamplesBsv(all_action[1:], all_action[:-1], AgentInfo(DistInfo(prob), torch.randn(T, B, device=dev, generator=g)),
                              torch.randn(1, B, device=dev, generator=g)),
        env=EnvSamples(obs, all_reward[1:], all_reward[:-1], torch.rand(T, B, device=dev, generator=g) < 0.002, None))


def ppo():
    from collections import namedtuple
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.samplers.collections import BatchSpec
    Spaces = namedtuple("Spaces", "observation action")
    agent = AtariFfAgent()
    agent.initialize(Spaces(namedtuple("O", "shape")(bench.IMAGE), namedtuple("Ac", "n")(bench.N_ACTIONS)))
    agent.to_device(0)
    T, B = bench.T_CFG, bench.B_CFG
    samples = resident_samples(T, B, bench.N_ACTIONS)
    algo = PPO(**bench.PPO_KW)
    algo.initialize(agent, 10 ** 6, BatchSpec(T, B), mid_batch_reset=True)
    agent.train_mode(0)
    algo.optimize_agent(0, samples)          # eager (lazy initialisation)
    algo.optimize_agent(1, samples)          # captures the minibatch graph
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    algo.optimize_agent(2, samples)          # 16 graph replays + the eager tail: ncu profiles the graphs' kernel nodes
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def gae():
    from rlpyt_b200.algos import utils as U
    for (T, B, algo) in [(128, 1 << 20, 1), (128, 256, 2), (128, 256, 1)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        r = torch.randn(T, B, device="cuda", generator=g)
        v = torch.randn(T, B, device="cuda", generator=g)
        d = torch.rand(T, B, device="cuda", generator=g) < 0.01
        b = torch.randn(1, B, device="cuda", generator=g)
        adv, ret = torch.empty_like(r), torch.empty_like(r)
        U.generalized_advantage_estimation(r, v, d, b, 0.99, 0.98, advantage_dest=adv, return_dest=ret, algo=algo)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        U.generalized_advantage_estimation(r, v, d, b, 0.99, 0.98, advantage_dest=adv, return_dest=ret, algo=algo)
        U.discount_return(r, d, b, 0.99, return_dest=ret, algo=algo)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


def replay():
    from tools.microbench import make_replay
    buf, Ex = make_replay(200_000)
    np.random.seed(0)
    pri = torch.rand(512, device="cuda") + 0.01
    buf.sample_batch(512)
    buf.update_batch_priorities(pri)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    buf.sample_batch(512)
    buf.update_batch_priorities(pri)
    buf.priority_tree.advance(128)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def conv():
    from rlpyt_b200.models.conv2_op import conv2_relu
    from rlpyt_b200.models import conv1_op
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.relu(torch.randn(8192, 16, 20, 20, device="cuda", generator=g))
    w = torch.randn(32, 16, 4, 4, device="cuda", generator=g) / 16
    b = torch.randn(32, device="cuda", generator=g)
    obs = torch.randint(0, 256, (8192, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    w1 = torch.randn(16, 4, 8, 8, device="cuda", generator=g) / 16
    b1 = torch.randn(16, device="cuda", generator=g)
    conv2_relu(x, w, b); conv1_op.conv1_u8_relu(w1, b1, obs, None)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    conv2_relu(x, w, b)
    conv1_op.conv1_u8_relu(w1, b1, obs, None)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def convbwd():
    """Backward of both conv layers: tcgen05 weight gradients (both) and the layer-2 input gradient."""
    from rlpyt_b200.models.conv2_op import conv2_relu
    from rlpyt_b200.models import conv1_op
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randint(0, 256, (8192, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    w1 = (torch.randn(16, 4, 8, 8, device="cuda", generator=g) / 16).requires_grad_(True)
    b1 = torch.randn(16, device="cuda", generator=g).requires_grad_(True)
    w = (torch.randn(32, 16, 4, 4, device="cuda", generator=g) / 16).requires_grad_(True)
    b = torch.randn(32, device="cuda", generator=g).requires_grad_(True)
    go = torch.randn(8192, 32, 10, 10, device="cuda", generator=g)
    for i in range(2):
        if i == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        y = conv2_relu(conv1_op.conv1_u8_relu(w1, b1, obs, None), w, b)
        y.backward(go)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    {"ppo": ppo, "gae": gae, "replay": replay, "conv": conv, "convbwd": convbwd}[sys.argv[1]]()
