#!/usr/bin/env python
"""Where does a PPO iteration / a sampler step spend its time?  torch.profiler (CUPTI) kernel table
for one optimize_agent call, and host-side phase timers for the sampler step loop.

    python tools/profile_step.py [--B 256] [--T 128] [--workers 32]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=128)
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--workers", type=int, default=32)
    ap.add_argument("--cudnn-benchmark", action="store_true")
    ap.add_argument("--channels-last", action="store_true")
    a = ap.parse_args()
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
    from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
    torch.backends.cudnn.benchmark = a.cudnn_benchmark
    sampler = GpuSampler(EnvCls=SyntheticAtariEnv, env_kwargs=bench.ENV_KW, batch_T=a.T, batch_B=a.B,
                         max_decorrelation_steps=0)
    agent = AtariFfAgent()
    sampler.initialize(agent, affinity=dict(cuda_idx=0, workers_cpus=[None] * a.workers, set_affinity=False),
                       seed=1, bootstrap_value=True)
    agent.to_device(0)
    if a.channels_last:
        agent.model.to(memory_format=torch.channels_last)
    algo = PPO(**bench.PPO_KW)
    algo.initialize(agent, 10 ** 6, sampler.batch_spec, mid_batch_reset=True)
    print("pinned step buffer:", sampler.host["pinned"], flush=True)
    try:
        samples, _ = sampler.obtain_samples(0)
        agent.train_mode(0)
        for i in range(2):
            algo.optimize_agent(i, samples)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        algo.optimize_agent(2, samples)
        torch.cuda.synchronize()
        print(f"optimize_agent: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            algo.optimize_agent(3, samples)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))

        # ---- sampler phases
        agent.sample_mode(4)
        ro, step_np = sampler.rollout, sampler.host["step_np"]
        for _ in range(2):
            sampler.obtain_samples(4)
        acc = dict(wait=0.0, upload=0.0, act=0.0, release=0.0)
        orig_upload, orig_act = ro.upload, ro.act

        def upload(*x, **k):
            t = time.perf_counter(); r = orig_upload(*x, **k); torch.cuda.synchronize(); acc["upload"] += time.perf_counter() - t; return r

        def act(*x, **k):
            t = time.perf_counter(); r = orig_act(*x, **k); acc["act"] += time.perf_counter() - t; return r
        ro.upload, ro.act = upload, act
        t0 = time.perf_counter()
        sampler.obtain_samples(5)
        tot = time.perf_counter() - t0
        print(f"obtain_samples: {1e3 * tot:.1f} ms total; upload(H2D, synced) {1e3 * acc['upload']:.1f} ms; "
              f"act(agent.step + D2H + sync) {1e3 * acc['act']:.1f} ms; rest (env workers + semaphores) "
              f"{1e3 * (tot - acc['upload'] - acc['act']):.1f} ms", flush=True)
        ro.upload, ro.act = orig_upload, orig_act
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for t in range(4):
                obs_dev = ro.upload(t, True)
                ro.act(t, obs_dev)
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=15, max_name_column_width=70))
    finally:
        try:
            sampler.shutdown()
        except Exception:
            os._exit(0)


if __name__ == "__main__":
    main()
