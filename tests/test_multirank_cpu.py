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
    a = make_affinity(0, 3, local_rank=0, ranks_per_node=1, node_share=0)       # no per-rank cap: 3 workers fit on 8 cpus
    assert a["set_affinity"] and len(a["workers_cpus"]) == 3 and a["cuda_idx"] == 0
    allowed = set(os.sched_getaffinity(0))
    assert all(set(c) <= allowed for c in a["workers_cpus"]) and set(a["master_cpus"]) <= allowed
    assert not set(a["master_cpus"]) & {c for w in a["workers_cpus"] for c in w}


def test_affinity_physical_cores_two_socket_box(monkeypatch):
    """The 2-socket B200 host (128 threads: cpu i and i+64 are hyper-threads of one core; GPUs 0-3 on cores 0-31,
    GPUs 4-7 on cores 32-63): every rank gets 8 whole cores at every N, no hardware thread is handed out twice,
    and two ranks never share a physical core (round 1 split THREADS and put ranks 0/2 on sibling threads)."""
    from rlpyt_b200.utils import affinity as A
    node = {0: list(range(0, 32)) + list(range(64, 96)), 1: list(range(32, 64)) + list(range(96, 128))}
    monkeypatch.setattr(A, "gpu_local_cpus", lambda idx: node[0 if idx < 4 else 1])
    monkeypatch.setattr(A, "_sibling_groups",
                        lambda cpus: [list(t) for t in sorted({tuple(sorted({c % 64, c % 64 + 64} & set(cpus))) for c in cpus})])
    monkeypatch.setattr(A.os, "sched_getaffinity", lambda pid: set(range(128)))
    for world in (1, 2, 4, 8):
        used_threads, used_cores = set(), set()
        for r in range(world):
            a = A.make_affinity(r, None, local_rank=r, ranks_per_node=world, node_share=8)
            threads = set(a["master_cpus"]) | {c for w in a["workers_cpus"] for c in w}
            cores = {t % 64 for t in threads}
            assert len(a["workers_cpus"]) == 7 and len(cores) == 8            # 1 master core + 7 worker cores
            assert not (threads & used_threads) and not (cores & used_cores)
            assert threads <= set(node[0 if r < 4 else 1])                    # GPU-local socket
            used_threads |= threads
            used_cores |= cores
        a = A.make_affinity(0, None, local_rank=0, ranks_per_node=world, node_share=8, smt_workers=True)
        assert len(a["workers_cpus"]) == 14 and len({w[0] for w in a["workers_cpus"]}) == 14


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def _async_worker(rank, world, port, out):
    """One rank of a two-rank asynchronous run on gloo: a fast and a slow sampler, stand-in algorithm whose
    ``optimize_agent`` all-reduces (as the real one's gradient all-reduce does)."""
    import time
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import test_async_runner_cpu as T
    from rlpyt_b200.runners.async_rl import AsyncRl
    from rlpyt_b200.utils.logging import TabularLogger

    class Algo(T.FakeAlgo):
        def optimize_agent(self, itr, samples=None, sampler_itr=None):
            t = torch.ones(1)
            dist.all_reduce(t)                                   # collective: every rank must make the same number of calls
            assert float(t) == world
            return super().optimize_agent(itr, samples, sampler_itr)

    class Agent(T.FakeAgent):
        device = torch.device("cpu")

        def data_parallel(self):
            self.dp = True

    sampler = T.FakeSampler(step_s=0.001 if rank == 0 else 0.004)   # rank 1 samples four times slower
    algo = Algo(min_steps_learn=0)
    runner = AsyncRl(algo=algo, agent=Agent(), sampler=sampler, n_steps=32 * 24, affinity=dict(cuda_idx=None), seed=1,
                     log_interval_steps=32 * 8, logger=TabularLogger(quiet=True))
    runner.throttle_wait = 0.002
    n_opt = runner.train()
    out[rank] = dict(n_opt=n_opt, calls=len(algo.calls), world=runner.world_size, sampler_itrs=len(sampler.calls))
    dist.barrier()
    dist.destroy_process_group()


def test_async_runner_ranks_agree_on_every_optimizer_iteration():
    """``AsyncRl`` under two ranks with unequal sampler speeds: the optimizer loops stay in lock-step (one tiny all-reduce
    of ready / quit flags before every iteration, the reference's ``opt_throttle`` barrier, async_rl.py:110-111) - a rank
    never enters the collective inside ``optimize_agent`` alone, and both stop together when the first sampler is done."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_async_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0]["world"] == out[1]["world"] == 2
    assert out[0]["n_opt"] == out[1]["n_opt"] == out[0]["calls"] == out[1]["calls"] > 0
    assert out[0]["sampler_itrs"] == 24                       # the fast rank finished its run; the slow one was cut short with it
