"""Parallel GPU sampler: env workers on CPU cores, ``agent.step`` on the B200, the ``[T,B]`` batch
resident in HBM (mirror of ``rlpyt/samplers/parallel/gpu/sampler.py:16-139`` +
``parallel/base.py:17-243`` + ``parallel/worker.py:37-101`` + ``gpu/action_server.py:17-74``).

Process model kept from the reference: one master (this process, one per GPU) and ``n_worker``
forked workers that each own a slice of the B environments; two barriers bracket a batch; one
(obs_ready, act_ready) semaphore pair per worker sequences every step (acquire/release counts are
identical to action_server.py:46-74 / gpu/collectors.py:25-48: T+1 obs releases, T act releases).

What changed is where the bytes live: the workers write only into the page-locked ``[B]`` step
buffer; the master DMA-copies each step's observations straight into ``observation[t]`` in HBM,
runs ``agent.step`` on resident slices and records action / prob / value on the device, so the
batch handed to the algorithm needs no further transfer (the reference re-uploads 925 MB per
iteration, ppo.py:72).
"""
import ctypes
import multiprocessing as mp
import os
import time

import numpy as np
import torch

from rlpyt_b200.samplers.base import BaseSampler
from rlpyt_b200.samplers.buffer import build_samples_buffer, pin_shared, unpin_shared, StepBuffer
from rlpyt_b200.samplers.collectors import GpuResetCollector
from rlpyt_b200.samplers.eval_collector import build_eval_collector
from rlpyt_b200.samplers.rollout import DeviceRollout
from rlpyt_b200.utils.collections import AttrDict
from rlpyt_b200.utils.seed import set_seed, set_envs_seeds
from rlpyt_b200.utils.synchronize import drain_queue, SpinSemaphore, SpinThenSleepSemaphore

_mp = mp.get_context("fork")


def sampling_process(common_kwargs, worker_kwargs):
    """Worker main loop (parallel/worker.py:37-101): build envs + collector, decorrelate, then
    collect a batch every time the master passes ``barrier_in`` until ``quit``."""
    # The child inherits every Python object of the master, CUDA handles included (graphs, streams, events of this and
    # earlier samplers).  Freeze them: a garbage collection in the child must never finalize a CUDA object in a
    # process that has no CUDA context (observed: the worker aborts inside cudaGraphExecDestroy, the master then waits
    # on the start-up barrier forever).
    import gc
    gc.freeze()
    c, w = AttrDict(**common_kwargs), AttrDict(**worker_kwargs)
    if w.cpus is not None:
        try:
            os.sched_setaffinity(0, w.cpus if isinstance(w.cpus, (list, tuple)) else [w.cpus])
        except (AttributeError, OSError):
            pass
    torch.set_num_threads(1)  # workers only run numpy/python (worker.py:25-28)
    if w.seed is not None:
        set_seed(w.seed)
    envs = [c.EnvCls(**c.env_kwargs) for _ in range(w.n_envs)]
    set_envs_seeds(envs, w.seed)
    collector = c.CollectorCls(rank=w.rank, envs=envs, env_info_np=w.env_info_np, batch_T=c.batch_T,
                               TrajInfoCls=c.TrajInfoCls, sync=w.sync, step_buffer_np=w.step_buffer_np,
                               global_B=c.global_B, env_ranks=w.env_ranks)
    traj_infos = collector.start_envs(c.max_decorrelation_steps)
    ctrl = c.ctrl
    ctrl.barrier_out.wait()
    while True:
        collector.reset_if_needed()
        ctrl.barrier_in.wait()
        if ctrl.quit.value:
            break
        traj_infos, completed = collector.collect_batch(traj_infos, ctrl.itr.value)
        for info in completed:
            c.traj_infos_queue.put(info)
        c.traj_infos_queue.put(None)                         # end of this worker's batch: the master drains up to one sentinel per worker
        ctrl.barrier_out.wait()
    for env in envs:
        env.close()


class GpuSampler(BaseSampler):

    gpu = True

    def __init__(self, *args, CollectorCls=GpuResetCollector, **kwargs):
        super().__init__(*args, CollectorCls=CollectorCls, **kwargs)

    # ------------------------------------------------------------------ API (samplers/base.py:49-63)
    def initialize(self, agent, affinity=None, seed=None, bootstrap_value=False, traj_info_kwargs=None,
                   world_size=1, rank=0):
        affinity = dict() if affinity is None else affinity
        B = self.batch_spec.B
        workers_cpus = affinity.get("workers_cpus")
        if workers_cpus is None:
            workers_cpus = [None] * min(B, max(1, (os.cpu_count() or 2) // max(1, world_size) - 1))
        n_envs_list = self._get_n_envs_list(len(workers_cpus))
        self.n_worker = n_worker = len(n_envs_list)
        self.world_size, self.rank = world_size, rank
        global_B = B * world_size
        env_ranks = list(range(rank * B, (rank + 1) * B))
        cuda_idx = affinity.get("cuda_idx", None)
        if cuda_idx is None:
            cuda_idx = torch.cuda.current_device()
        self.device = torch.device("cuda", cuda_idx)

        from rlpyt_b200 import _lib
        _lib.load()  # before the fork: workers use its host-side streaming copy
        env = self.EnvCls(**self.env_kwargs)
        if not getattr(self, "_agent_preinitialized", False):   # the asynchronous samplers initialize the agent in async_initialize
            agent.initialize(env.spaces, share_memory=False, global_B=global_B, env_ranks=env_ranks)
        self.agent = agent
        self.samples, self.host, examples = build_samples_buffer(
            agent, env, self.batch_spec, bootstrap_value, device=self.device, share_host=True)
        env.close()
        del env

        self.ctrl = AttrDict(
            quit=_mp.RawValue(ctypes.c_bool, False),
            barrier_in=_mp.Barrier(n_worker + 1),
            barrier_out=_mp.Barrier(n_worker + 1),
            itr=_mp.RawValue(ctypes.c_long, 0),
        )
        self.traj_infos_queue = _mp.Queue()
        # step-loop handshakes (RLPYT_B200_SAMPLER_SYNC): "futex" = multiprocessing semaphores as in the
        # reference (default); "spin" = spinning single-producer/single-consumer counters; "hybrid" = the
        # same semaphores behind a bounded spin (utils/synchronize.py).  On the build host a 7-worker
        # round trip costs 330 us / 9 us / 19 us; the non-default modes stay opt-in until they have been
        # measured on the GPU box - tests/test_sampler_protocol_cpu.py runs the
        # real worker loop against all three.
        mode = os.environ.get("RLPYT_B200_SAMPLER_SYNC", "futex")
        if mode == "spin":
            make_sem = lambda: SpinSemaphore(_mp)
        elif mode == "hybrid":
            make_sem = lambda: SpinThenSleepSemaphore(_mp)
        elif mode == "futex":
            make_sem = lambda: _mp.Semaphore(0)
        else:
            raise ValueError(f"RLPYT_B200_SAMPLER_SYNC must be futex, spin or hybrid (got {mode!r})")
        self.sync = AttrDict(obs_ready=[make_sem() for _ in range(n_worker)],
                             act_ready=[make_sem() for _ in range(n_worker)])
        if traj_info_kwargs:
            for k, v in traj_info_kwargs.items():
                setattr(self.TrajInfoCls, "_" + k, v)

        common = dict(EnvCls=self.EnvCls, env_kwargs=self.env_kwargs, batch_T=self.batch_spec.T,
                      CollectorCls=self.CollectorCls, TrajInfoCls=self.TrajInfoCls,
                      traj_infos_queue=self.traj_infos_queue, ctrl=self.ctrl,
                      max_decorrelation_steps=self.max_decorrelation_steps, global_B=global_B)
        step_np, env_info_np = self.host["step_np"], self.host["env_info_np"]
        self.workers, i_env, g_env = [], 0, B * rank
        self.worker_slices = []
        for w_rank, n_envs in enumerate(n_envs_list):
            sl = slice(i_env, i_env + n_envs)
            self.worker_slices.append(sl)
            wk = dict(rank=w_rank, env_ranks=list(range(g_env, g_env + n_envs)),
                      seed=None if seed is None else seed + w_rank,
                      cpus=workers_cpus[w_rank] if affinity.get("set_affinity", True) else None,
                      n_envs=n_envs, step_buffer_np=step_np[sl],
                      env_info_np=None if env_info_np is None else env_info_np[:, sl],
                      sync=AttrDict(obs_ready=self.sync.obs_ready[w_rank], act_ready=self.sync.act_ready[w_rank]))
            i_env += n_envs
            g_env += n_envs
            self.workers.append(_mp.Process(target=sampling_process,
                                            kwargs=dict(common_kwargs=common, worker_kwargs=wk), daemon=True))
        import gc
        gc.collect()                                         # drop cyclic garbage (old CUDA graphs ...) here, where CUDA is valid
        for w in self.workers:
            w.start()
        if not self.host["pinned"]:  # page-lock after the fork so the children never see CUDA state
            self.host["pinned"] = all(pin_shared(a) for a in step_np)
        if affinity.get("set_affinity", True) and affinity.get("master_cpus"):
            try:
                os.sched_setaffinity(0, affinity["master_cpus"])
            except (AttributeError, OSError):
                pass
        try:   # workers decorrelated, first observations are in the step buffer
            self.ctrl.barrier_out.wait(timeout=300)
        except Exception as e:  # noqa: BLE001 - a worker that died during start-up must not hang the job
            dead = [i for i, w in enumerate(self.workers) if not w.is_alive()]
            for w in self.workers:
                if w.is_alive():
                    w.terminate()
            raise RuntimeError(f"sampler workers did not come up (dead: {dead})") from e
        self.rollout = DeviceRollout(self.samples, self.host, agent, self.device)
        self.rollout.set_worker_chunks(self.worker_slices)
        self.profile = dict(wait_envs_s=0.0, device_step_s=0.0, release_s=0.0, steps=0)
        self.rollout.in_action.copy_(self.host["step_pyt"].action)
        self.samples_pyt = self.samples
        self.eval_collector = build_eval_collector(self, agent, seed)     # after the fork: master-only envs
        return examples

    def obtain_samples(self, itr):
        """gpu/sampler.py:45-56.  Returns the SAME preallocated device-resident ``Samples`` every call."""
        self.agent.sample_mode(itr)
        self.ctrl.itr.value = itr
        self.ctrl.barrier_in.wait()
        self.serve_actions(itr)
        self.ctrl.barrier_out.wait()
        # one sentinel per worker marks the end of its puts for this batch: every trajectory that ended in the batch is
        # returned WITH the batch (the reference's non-blocking drain, parallel/base.py:69, can miss an item whose
        # queue feeder thread has not pushed it yet and report it one batch late)
        traj_infos = drain_queue(self.traj_infos_queue, n_sentinel=self.n_worker)
        return self.samples, traj_infos

    def evaluate_agent(self, itr):
        """Evaluation runs in the master process (samplers/eval_collector.py): the training step loop
        and its worker protocol stay untouched."""
        if self.eval_collector is None:
            raise RuntimeError("evaluate_agent needs eval_n_envs > 0 (and eval_max_steps) at construction")
        return self.eval_collector.collect_evaluation(itr)

    def shutdown(self):
        """Ask the workers to leave at the next batch boundary.  If the master died in the middle of a batch the workers
        are parked on their action semaphores and will never reach the barrier: time out and terminate them instead of
        hanging the caller (a failed run must cost seconds, not the job's wall-clock limit)."""
        import threading
        self.ctrl.quit.value = True
        try:
            self.ctrl.barrier_in.wait(timeout=10)
        except threading.BrokenBarrierError:
            pass
        for w in self.workers:
            w.join(timeout=5)
            if w.is_alive():
                w.terminate()
                w.join(timeout=5)
        torch.cuda.synchronize(self.device)
        if self.host.get("pinned"):
            for a in self.host["step_np"]:
                unpin_shared(a)
            self.host["pinned"] = False

    # ------------------------------------------------------------------ action server
    def serve_actions(self, itr):
        """Master half of the step loop (gpu/action_server.py:17-74)."""
        obs_ready, act_ready = self.sync.obs_ready, self.sync.act_ready
        step_np, ro = self.host["step_np"], self.rollout
        T = self.batch_spec.T
        wait_reset = not self.mid_batch_reset
        prof = self.profile if os.environ.get("RLPYT_B200_SAMPLER_PROFILE") == "1" else None
        clock = time.perf_counter
        n_worker = len(obs_ready)
        chunked = os.environ.get("RLPYT_B200_SAMPLER_CHUNKED", "0") == "1"   # 1: upload each worker's rows as it signals (see AlternatingSampler)
        for t in range(T):
            t0 = clock() if prof is not None else 0.0
            # workers wrote obs(t), reward(t-1), done(t-1): each worker's rows go to HBM as soon as it has signalled,
            # overlapping the H2D with the workers that are still stepping (action_server.py:46-48 waits for all first)
            pending = list(range(n_worker))
            while pending:
                progressed = False
                for w in list(pending):
                    if obs_ready[w].acquire(block=False):
                        if chunked:
                            ro.upload_worker_rows(t, w)
                        pending.remove(w)
                        progressed = True
                if pending and not progressed:
                    w = pending.pop(0)
                    obs_ready[w].acquire()
                    if chunked:
                        ro.upload_worker_rows(t, w)
            t1 = clock() if prof is not None else 0.0
            done_now = step_np.done
            if self.mid_batch_reset and np.any(done_now):
                for b in np.where(done_now)[0]:
                    self.agent.reset_one(idx=b)
            ro.step(t, zero_inputs_on_done=True, blank_done_rows=wait_reset, obs_done=chunked)
            t2 = clock() if prof is not None else 0.0
            for s in act_ready:
                s.release()
            if prof is not None:   # where a step's wall time goes on the master: waiting for the envs / device / wake-ups
                prof["wait_envs_s"] += t1 - t0
                prof["device_step_s"] += t2 - t1
                prof["release_s"] += clock() - t2
                prof["steps"] += 1
        for s in obs_ready:
            s.acquire()
            assert not s.acquire(block=False)  # drained (action_server.py:63)
        ro.finish()
        # finish() only ENQUEUES the DMA of reward(T-1)/done(T-1) out of the pinned step buffer: the host
        # must not touch that buffer (the zeroing below) before the copies have been executed.
        torch.cuda.current_stream(self.device).synchronize()
        if np.any(step_np.done):  # reset at end of batch; ready for the next (action_server.py:67-71)
            ended = np.where(step_np.done)[0]
            step_np.action[ended] = 0
            step_np.reward[ended] = 0
            for b in ended:
                self.agent.reset_one(idx=b)
            ro.zero_inputs_where_done()
        torch.cuda.current_stream(self.device).synchronize()
        ro.end_batch()
        for s in act_ready:
            assert not s.acquire(block=False)

    # ------------------------------------------------------------------ helpers
    def _get_n_envs_list(self, n_worker):
        """parallel/base.py:222-243."""
        B = self.batch_spec.B
        n_worker = min(n_worker, B)
        n_envs_list = [B // n_worker] * n_worker
        for b in range(B % n_worker):
            n_envs_list[b] += 1
        return n_envs_list
