"""Two ranks on two GPUs (NCCL): after K PPO updates on DIFFERENT per-rank samples the data-parallel replicas hold
bit-identical parameters (one flat-gradient all-reduce per update, mean folded into the fused clip + Adam kernel:
rlpyt_b200/algos/optim.py; the reference relies on DistributedDataParallel, rlpyt/agents/base.py:118-136) - and they differ
from what a single rank would have learnt from its own samples alone.  Skips unless two GPUs are visible."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    import numpy as np
    from collections import namedtuple
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    from rlpyt_b200.agents.pg.base import AgentInfo
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.distributions.categorical import DistInfo
    from rlpyt_b200.samplers.collections import AgentSamplesBsv, BatchSpec, EnvSamples, Samples
    T, B, A, image = 16, 32, 6, (4, 84, 84)
    dev = torch.device("cuda", rank)
    torch.manual_seed(100 * rank)                               # different initial weights per rank (sync_rl.py:82)
    Spaces = namedtuple("Spaces", "observation action")
    agent = AtariFfAgent()
    agent.initialize(Spaces(namedtuple("O", "shape")(image), namedtuple("Ac", "n")(A)))
    agent.to_device(rank)
    agent.data_parallel()                                       # rank 0's parameters everywhere
    algo = PPO(gae_lambda=0.95, minibatches=2, epochs=2)
    algo.initialize(agent, 10, BatchSpec(T, B), mid_batch_reset=True, world_size=world, rank=rank)
    g = torch.Generator(device=dev).manual_seed(7 + rank)       # different samples per rank
    all_a = torch.randint(0, A, (T + 1, B), device=dev, generator=g)
    all_r = torch.randn(T + 1, B, device=dev, generator=g)
    samples = Samples(
        agent=AgentSamplesBsv(action=all_a[1:], prev_action=all_a[:-1],
                              agent_info=AgentInfo(dist_info=DistInfo(prob=torch.softmax(torch.randn(T, B, A, device=dev, generator=g), -1)),
                                                   value=torch.randn(T, B, device=dev, generator=g)),
                              bootstrap_value=torch.randn(1, B, device=dev, generator=g)),
        env=EnvSamples(observation=torch.randint(0, 256, (T, B) + image, dtype=torch.uint8, device=dev, generator=g),
                       reward=all_r[1:], prev_reward=all_r[:-1], done=torch.rand(T, B, device=dev, generator=g) < 0.05, env_info=None))
    np.random.seed(3)                                           # same shuffles on both ranks (not required, keeps the run reproducible)
    agent.train_mode(0)
    w0 = algo.optimizer.flat_param.clone()
    for itr in range(3):                                        # eager iteration, graph capture, graph replay
        algo.optimize_agent(itr, samples)
    flat = algo.optimizer.flat_param
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    out[rank] = dict(identical=all(torch.equal(gathered[0], x) for x in gathered), moved=bool((flat != w0).any()),
                     finite=bool(torch.isfinite(flat).all()), updates=algo.update_counter)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_hold_identical_parameters_after_ppo_updates():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        assert out[r]["identical"] and out[r]["moved"] and out[r]["finite"] and out[r]["updates"] == 12
