"""Multi-process host logic on CPU (gloo, world_size 2, 127.0.0.1): the pieces of the data-parallel
path that do not need a GPU - parameter replication of ``agent.data_parallel()``, per-rank
seeding / env ranks, CPU placement, and the rank-0-only contract of ``bench.py --impl reference``."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from collections import namedtuple
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    torch.manual_seed(100 * rank)  # different init per rank (sync_rl.py:82 seeds ranks differently)
    agent = AtariFfAgent()
    Spaces = namedtuple("Spaces", "observation action")
    agent.initialize(Spaces(namedtuple("O", "shape")((4, 36, 36)), namedtuple("A", "n")(5)))
    before = torch.cat([p.detach().reshape(-1) for p in agent.parameters()]).clone()
    agent.data_parallel()
    after = torch.cat([p.detach().reshape(-1) for p in agent.parameters()])
    gathered = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    # the flat-gradient mean: SUM all-reduce then 1/world, as FlatAdam.clip_and_step does
    g = torch.full((7,), float(rank + 1))
    dist.all_reduce(g)
    out[rank] = dict(changed=bool((before != after).any()), same=all(torch.equal(gathered[0], x) for x in gathered),
                     world=agent.world_size, mean=float((g / world)[0]))
    dist.destroy_process_group()


def test_data_parallel_replicates_rank0_parameters():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0]["same"] and out[1]["same"]
    assert out[0]["changed"] is False and out[1]["changed"] is True   # rank 1 adopted rank 0's weights
    assert out[0]["world"] == out[1]["world"] == 2
    assert out[0]["mean"] == out[1]["mean"] == 1.5


def test_affinity_split_is_disjoint_per_rank():
    from rlpyt_b200.utils.affinity import make_affinity
    a = make_affinity(0, 3, local_rank=0, ranks_per_node=1)
    assert a["set_affinity"] and len(a["workers_cpus"]) == 3 and a["cuda_idx"] == 0
    allowed = set(os.sched_getaffinity(0))
    assert all(set(c) <= allowed for c in a["workers_cpus"]) and set(a["master_cpus"]) <= allowed


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""
