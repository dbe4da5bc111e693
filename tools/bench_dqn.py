#!/usr/bin/env python
"""DQN learner benchmark (SURVEY.md 8(d) config 4 + 8(f) row 1): prioritized frame replay in HBM,
Double-DQN, n-step 3 - one "update" = sample_batch -> two/three network forwards -> fused loss ->
backward -> clip+Adam -> priority update.  Not part of the driver's bench contract (bench.py keeps
BASELINE.json's PPO metric); written at the end of round 1 after the GPU budget was spent, so it has
NOT been measured yet - run it first thing next round:

    python tools/bench_dqn.py [--frames 1000000] [--batch 512] [--updates 200] [--cpu-updates 3]

Prints one JSON line: updates/s and sampled transitions/s on the GPU (CUDA events, replay larger than
L2), the same update on the host (oracle replay + torch-CPU network + oracle loss; bounded sample),
and the three phases of an update (sample / learn / priorities) timed separately.
"""
import argparse
import json
import os
import sys
import time
from collections import namedtuple

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

IMG, A, T, B = (4, 84, 84), 6, 128, 256
Spaces = namedtuple("Spaces", "observation action")


def synth(seed, dev):
    from rlpyt_b200.agents.dqn.dqn_agent import AgentInfo
    from rlpyt_b200.samplers.collections import AgentSamples, EnvSamples, Samples
    g = torch.Generator(device=dev).manual_seed(seed)
    obs = torch.randint(0, 256, (T, B) + IMG, dtype=torch.uint8, device=dev, generator=g)
    act = torch.randint(0, A, (T + 1, B), device=dev, generator=g)
    rew = torch.randn(T + 1, B, device=dev, generator=g)
    done = torch.rand(T, B, device=dev, generator=g) < 0.005
    return Samples(agent=AgentSamples(act[1:], act[:-1], AgentInfo(q=torch.zeros(T, B, A, device=dev))),
                   env=EnvSamples(obs, rew[1:], rew[:-1], done, None))


def events(fn, reps):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--updates", type=int, default=200)
    ap.add_argument("--cpu-updates", type=int, default=3)
    args = ap.parse_args()
    from rlpyt_b200.agents.dqn.atari.atari_dqn_agent import AtariDqnAgent
    from rlpyt_b200.algos.dqn.dqn import DQN
    from rlpyt_b200.samplers.collections import BatchSpec
    from rlpyt_b200 import _lib
    torch.manual_seed(0)
    np.random.seed(0)
    agent = AtariDqnAgent()
    agent.initialize(Spaces(namedtuple("O", "shape")(IMG), namedtuple("Ac", "n")(A)))
    agent.to_device(0)
    algo = DQN(batch_size=args.batch, min_steps_learn=0, replay_size=args.frames, replay_ratio=8, n_step_return=3,
               double_dqn=True, prioritized_replay=True, target_update_interval=312)
    examples = dict(observation=np.zeros(IMG, np.uint8), action=np.int64(0), reward=np.float32(0),
                    done=np.bool_(False))
    algo.initialize(agent, n_itr=10 ** 6, batch_spec=BatchSpec(T, B), mid_batch_reset=True, examples=examples)
    buf = algo.replay_buffer
    data = [synth(s, "cuda") for s in range(2)]
    for i in range(buf.T // T + 2):                       # fill past one wrap
        buf.append_samples(algo.samples_to_buffer(data[i % 2]))
    agent.train_mode(0)

    def update():
        batch = buf.sample_batch(args.batch)
        algo.optimizer.zero_grad()
        loss, td = algo.loss(batch)
        loss.backward()
        algo.optimizer.clip_and_step(algo.clip_grad_norm)
        buf.update_batch_priorities(td)

    for _ in range(5):
        update()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.updates):
        update()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / args.updates
    # phases
    t_sample = events(lambda: buf.sample_batch(args.batch), 20)
    batch = buf.sample_batch(args.batch)

    def learn():
        algo.optimizer.zero_grad()
        loss, td = algo.loss(batch)
        loss.backward()
        algo.optimizer.clip_and_step(algo.clip_grad_norm)
        return td
    t_learn = events(learn, 20)
    td = learn()
    t_pri = events(lambda: buf.update_batch_priorities(td), 20)
    out = {"metric": "DQN updates/s (prioritized frame replay, Double-DQN, n-step 3)", "value": 1.0 / dt,
           "unit": "updates/s", "transitions_per_s": args.batch / dt, "ms_per_update": dt * 1e3,
           "phases_ms": {"sample_batch": t_sample * 1e3, "forward_loss_backward_step": t_learn * 1e3,
                         "update_priorities": t_pri * 1e3},
           "config": {"workload": f"replay {buf.size} frames x (84,84) u8, B={B}, batch {args.batch}, AtariDqnModel A={A}",
                      "l2": "replay store (7 GB) larger than L2"},
           "gpu_launches_per_update": (_lib.launch_count - launches0) / args.updates, "dtype": "f32",
           "data": "synthetic"}
    if args.cpu_updates > 0:
        out["cpu_baseline"] = cpu_baseline(agent, args)
    print(json.dumps(out), flush=True)


def cpu_baseline(agent, args):
    """The same update on the host: oracle FrameReplay (numpy, the reference's algorithm), the reference
    network on torch-CPU, oracle loss, torch Adam.  Bounded: a 64 K-frame replay (sampling cost per batch
    does not depend on the replay size beyond tree depth) and a few updates."""
    from oracle.dqn_loss import dqn_loss
    from oracle.replay import FrameReplay
    from rlpyt_b200.models.dqn.atari_dqn_model import AtariDqnModel
    rep = FrameReplay(IMG, 65_536, B, discount=0.99, n_step_return=3, prioritized=True)
    rng = np.random.default_rng(0)
    s = dict(observation=rng.integers(0, 256, (T, B) + IMG, dtype=np.uint8), action=rng.integers(0, A, (T, B)),
             reward=rng.standard_normal((T, B)).astype(np.float32), done=rng.random((T, B)) < 0.005)
    for _ in range(3):
        rep.append_samples(s)
    model, target = AtariDqnModel(IMG, A), AtariDqnModel(IMG, A)
    sd = {k: v.detach().cpu() for k, v in agent.model.state_dict().items()}
    model.load_state_dict(sd)
    target.load_state_dict(sd)
    opt = torch.optim.Adam(model.parameters(), lr=2.5e-4, eps=0.01 / args.batch)
    t = torch.from_numpy

    def update():
        b = rep.sample_batch(args.batch)
        opt.zero_grad()
        qs = model(t(b["observation"]), None, None)
        with torch.no_grad():
            tq = target(t(b["target_observation"]), None, None)
            nq = model(t(b["target_observation"]), None, None)
        _loss, td, grad = dqn_loss(qs.detach(), tq, nq, t(b["action"]), t(b["return_"]), t(b["done_n"]),
                                   t(b["is_weights"]), 0.99, 3, 1.0)
        qs.backward(grad)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        rep.update_batch_priorities(td.numpy())

    update()
    t0 = time.perf_counter()
    for _ in range(args.cpu_updates):
        update()
    dt = (time.perf_counter() - t0) / args.cpu_updates
    return {"value": 1.0 / dt, "unit": "updates/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{args.cpu_updates} updates, batch {args.batch}, 64 K-frame oracle replay, torch-CPU network"}


if __name__ == "__main__":
    main()
