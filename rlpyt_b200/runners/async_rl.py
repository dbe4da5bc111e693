"""Asynchronous runner: sampling and optimization run concurrently (mirror of ``rlpyt/runners/async_rl.py:21-612``
``AsyncRlBase`` / ``AsyncRl`` / ``AsyncRlEval`` with ``run_async_sampler`` / ``run_async_sampler_eval`` /
``memory_copier``).  Same constructor, ``train()`` loop, throttle arithmetic
(``replay_ratio`` bounds optimizer speed against sampler speed), double-buffer hand-shake (``sample_ready`` /
``sample_copied`` semaphore pairs), logging rows and shutdown order.

B200 design.  The reference needs four kinds of PROCESSES (sampler, two memory copiers, optimizer[s]) because its
sample batches and its replay buffer are numpy arrays in OS shared memory and its sampler may own other GPUs.  On this
path everything that touches sample data lives in the HBM of the one GPU the process owns (one process per GPU, as
everywhere in this package), so the same roles are THREADS of that process, each issuing to its own CUDA stream:

    sampler thread   obtain_samples(itr, db_idx): CPU env workers (forked processes, as in synchronous mode) ->
                     agent.step graphs on the sampler's stream(s) -> device-to-device publish into double_buffer[db_idx]
    copier thread    replay_buffer.append_samples(algo.samples_to_buffer(double_buffer[i])) on its own stream: frame
                     ring, n-step returns and sum-tree advance are kernels (rlpyt_b200.replays), microseconds per batch
    main thread      the optimizer loop below: algo.optimize_agent(itr, sampler_itr=...) + agent.send_shared_memory()

Host-side ordering is the reference's (semaphores, the replay's read-write lock); DEVICE ordering between the streams
is by CUDA events (``AsyncSamplerMixin.acquire_batch / release_batch``, ``replays.async_.StreamFence``) - no thread
ever blocks on the GPU for another thread's work.  Python threads share the interpreter lock: the sampler's master loop
spends its time in semaphore waits and stream synchronisation (both release it), and ``sys.setswitchinterval`` is
lowered for the run so that a step's few microseconds of Python are not held up by the optimizer's launch loop.

Multi-GPU (``torchrun``, one process per GPU, NCCL): every rank runs this runner with its own sampler and replay; the
gradient all-reduce inside ``optimize_agent`` is collective, so the ranks agree before every optimizer iteration on
"everybody has enough new samples" / "somebody is done" with one tiny all-reduce (the reference's ``opt_throttle``
barrier, async_rl.py:110-111, 484-488).
"""
import math
import sys
import threading
import time
from collections import deque

import torch

from rlpyt_b200.utils.collections import AttrDict
from rlpyt_b200.utils.seed import make_seed, set_seed
from rlpyt_b200.utils.synchronize import drain_queue

THROTTLE_WAIT = 0.05


class _Value:
    """``mp.Value``-shaped cell for thread use (``.value``, ``get_lock()``)."""

    def __init__(self, value=0):
        self.value = value
        self._lock = threading.RLock()

    def get_lock(self):
        return self._lock


class AsyncRlBase:

    _eval = False

    def __init__(self, algo, agent, sampler, n_steps, affinity=None, seed=None, log_interval_steps=1e5, logger=None):
        self.algo, self.agent, self.sampler = algo, agent, sampler
        self.n_steps = int(n_steps)
        self.affinity = affinity
        self.seed = seed
        self.log_interval_steps = int(log_interval_steps)
        if logger is None:
            from rlpyt_b200.utils.logging import TabularLogger
            logger = TabularLogger()
        self.logger = logger
        self.throttle_wait = THROTTLE_WAIT

    # ------------------------------------------------------------------ the optimizer loop (async_rl.py:78-132)
    def train(self):
        logger = self.logger
        old_switch = sys.getswitchinterval()
        sys.setswitchinterval(min(old_switch, 2e-4))
        try:
            throttle_itr, delta_throttle_itr = self.startup()
            throttle_time = 0.
            sampler_itr = itr = 0
            if self._eval:
                while self.ctrl.sampler_itr.value < 1 and not self._worker_failed():   # the sampler evaluates first
                    time.sleep(self.throttle_wait)
                traj_infos = drain_queue(self.traj_infos_queue, n_sentinel=1)
                self.store_diagnostics(0, 0, traj_infos, ())
                self.log_diagnostics(0, 0, 0)
            log_counter = 0
            while True:                                   # until the sampler reaches n_steps and sets ctrl.quit
                logger.set_iteration(itr)
                with logger.prefix(f"opt_itr #{itr} "):
                    while True:
                        ready = self.ctrl.sampler_itr.value >= throttle_itr
                        quit_ = bool(self.ctrl.quit.value) or self._worker_failed()
                        ready, quit_ = self._agree(ready, quit_)
                        if ready or quit_:
                            break
                        time.sleep(self.throttle_wait)
                        throttle_time += self.throttle_wait
                    if quit_:
                        break
                    throttle_itr += delta_throttle_itr
                    opt_info = self.algo.optimize_agent(itr, sampler_itr=self.ctrl.sampler_itr.value)
                    self.agent.send_shared_memory()       # to the sampler (staging copy in HBM)
                    sampler_itr = self.ctrl.sampler_itr.value
                    traj_infos = list() if self._eval else drain_queue(self.traj_infos_queue)
                    self.store_diagnostics(itr, sampler_itr, traj_infos, opt_info)
                    if sampler_itr // self.log_interval_itrs > log_counter:
                        if self._eval:
                            with self.ctrl.sampler_itr.get_lock():
                                traj_infos = drain_queue(self.traj_infos_queue, n_sentinel=1)
                            self.store_diagnostics(itr, sampler_itr, traj_infos, ())
                        self.log_diagnostics(itr, sampler_itr, throttle_time)
                        log_counter += 1
                        throttle_time = 0.
                itr += 1
            sampler_itr = self.ctrl.sampler_itr.value
            traj_infos = drain_queue(self.traj_infos_queue)
            if traj_infos or not self._eval:
                self.store_diagnostics(itr, sampler_itr, traj_infos, ())
                self.log_diagnostics(itr, sampler_itr, throttle_time)
        finally:
            self.shutdown()
            sys.setswitchinterval(old_switch)
        if self._errors:
            raise RuntimeError(f"asynchronous {self._errors[0][0]} thread failed") from self._errors[0][1]
        return itr

    def _agree(self, ready, quit_):
        """All ranks take the same branch (the gradient all-reduce is collective): ready = everybody is, quit = anybody."""
        if self.world_size == 1:
            return ready, quit_
        import torch.distributed as dist
        flags = torch.tensor([1 if ready else 0, 0 if quit_ else 1], dtype=torch.int32, device=self.agent.device)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        r, nq = flags.tolist()
        return bool(r), not bool(nq)

    def _worker_failed(self):
        return bool(self._errors)

    # ------------------------------------------------------------------ start-up (async_rl.py:134-186)
    def startup(self):
        logger = self.logger
        self._errors = []
        if self.seed is None:
            self.seed = make_seed()
        set_seed(self.seed)
        aff = self._affinities()
        cuda_idx = aff.optimizer.get("cuda_idx", None)
        if cuda_idx is None:
            cuda_idx = torch.cuda.current_device() if torch.cuda.is_available() else None
        self.world_size = 1
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.world_size = dist.get_world_size()
        except Exception:  # noqa: BLE001
            pass
        async_kwargs = dict(agent=self.agent, bootstrap_value=getattr(self.algo, "bootstrap_value", False),
                            traj_info_kwargs=self.get_traj_info_kwargs(), seed=self.seed)
        if cuda_idx is not None:
            async_kwargs["device"] = torch.device("cuda", cuda_idx)
        double_buffer, examples = self.sampler.async_initialize(**async_kwargs)
        self.sampler_batch_size = self.sampler.batch_spec.size
        n_itr = self.get_n_itr()
        # optimizer side first: the agent goes to the device, then the replay is allocated there (the reference
        # allocates the shared-memory replay before it forks; here the order only has to put the ring into HBM)
        self.agent.to_device(cuda_idx)
        if self.world_size > 1:
            self.agent.data_parallel()
        replay_buffer = self.algo.async_initialize(agent=self.agent, sampler_n_itr=n_itr,
                                                   batch_spec=self.sampler.batch_spec,
                                                   mid_batch_reset=self.sampler.mid_batch_reset, examples=examples,
                                                   world_size=self.world_size)
        self.algo.optim_initialize(rank=getattr(self, "rank", 0))
        throttle_itr = 1 + getattr(self.algo, "min_steps_learn", 0) // self.sampler_batch_size
        delta_throttle_itr = (self.algo.batch_size * self.world_size * self.algo.updates_per_optimize /
                              (self.sampler_batch_size * self.algo.replay_ratio))
        # sampler: its own parameter copy, env workers forked while this process is still single-threaded
        self.sampler.initialize(aff.sampler)
        self.initialize_logging()
        self.launch_workers(n_itr, double_buffer, replay_buffer)
        logger.log(f"Asynchronous runner: {n_itr} sampler iterations, optimizer may start at sampler itr {throttle_itr}, "
                   f"then one optimize_agent per {delta_throttle_itr:.3g} sampler iterations.")
        return throttle_itr, delta_throttle_itr

    def _affinities(self):
        """``affinity`` may be the reference's structure with ``.sampler`` / ``.optimizer`` (list, first entry used:
        one process drives one GPU), a dict with those keys, or one flat dict used for both."""
        a = self.affinity if self.affinity is not None else dict()
        smp = a.get("sampler", None) if isinstance(a, dict) else getattr(a, "sampler", None)
        opt = a.get("optimizer", None) if isinstance(a, dict) else getattr(a, "optimizer", None)
        if smp is None and opt is None:
            smp = opt = a
        if isinstance(opt, (list, tuple)):
            opt = opt[0]
        return AttrDict(sampler=dict(smp or {}), optimizer=dict(opt or {}))

    def get_n_itr(self):
        """async_rl.py:188-196."""
        log_interval_itrs = max(self.log_interval_steps // self.sampler_batch_size, 1)
        n_itr = math.ceil(self.n_steps / self.log_interval_steps) * log_interval_itrs
        self.log_interval_itrs = log_interval_itrs
        self.n_itr = n_itr
        self.logger.log(f"Running {n_itr} sampler iterations.")
        return n_itr

    def build_ctrl(self):
        """async_rl.py:198-214 with thread primitives."""
        return AttrDict(
            quit=_Value(False),
            sample_ready=[threading.Semaphore(0) for _ in range(2)],     # double buffer
            sample_copied=[threading.Semaphore(1) for _ in range(2)],
            sampler_itr=_Value(0),
            published=_Value(0),            # batches the sampler has completed (the copier tells them from the quit sentinel)
            eval_time=_Value(0.),
        )

    def launch_workers(self, n_itr, double_buffer, replay_buffer):
        import queue
        self.traj_infos_queue = queue.Queue()
        self.ctrl = self.build_ctrl()
        device = getattr(self.sampler, "device", None)
        target = run_async_sampler_eval if self._eval else run_async_sampler
        kwargs = dict(sampler=self.sampler, ctrl=self.ctrl, traj_infos_queue=self.traj_infos_queue, n_itr=n_itr,
                      device=device, errors=self._errors, logger=self.logger)
        if self._eval:
            kwargs["eval_itrs"] = self.log_interval_itrs
        self.sampler_thread = threading.Thread(target=target, kwargs=kwargs, name="async-sampler", daemon=True)
        self.memcpy_thread = threading.Thread(
            target=memory_copier, name="async-memcpy", daemon=True,
            kwargs=dict(sampler=self.sampler, samples_to_buffer=self.algo.samples_to_buffer, replay_buffer=replay_buffer,
                        ctrl=self.ctrl, device=device, errors=self._errors, logger=self.logger))
        self.memcpy_thread.start()
        self.sampler_thread.start()

    def shutdown(self):
        """async_rl.py:272-286."""
        ctrl = getattr(self, "ctrl", None)
        if ctrl is None:
            return
        if hasattr(self, "pbar"):
            self.pbar = None
        self.logger.log("Master optimizer shutting down, joining sampler thread...")
        ctrl.quit.value = True                                # also stops a sampler that is still running (error paths)
        for s in ctrl.sample_copied:
            s.release()
        self.sampler_thread.join(timeout=60)
        for s in ctrl.sample_ready:
            s.release()
        self.logger.log("Joining memory copier...")
        self.memcpy_thread.join(timeout=30)
        if not getattr(self.sampler, "_shut_down", False):    # the sampler thread shuts it down itself on a clean exit
            try:
                self.sampler.shutdown()
            except Exception:  # noqa: BLE001
                pass
        self.logger.log("All threads joined.  Training complete.")

    # ------------------------------------------------------------------ logging (async_rl.py:288-370)
    def initialize_logging(self):
        self._opt_infos = {k: list() for k in self.algo.opt_info_fields}
        self._start_time = self._last_time = time.time()
        self._last_itr = 0
        self._last_sampler_itr = 0
        self._last_update_counter = 0

    def get_itr_snapshot(self, itr, sampler_itr):
        return dict(itr=itr, sampler_itr=sampler_itr, cum_steps=sampler_itr * self.sampler_batch_size,
                    cum_updates=self.algo.update_counter, agent_state_dict=self.agent.state_dict(),
                    optimizer_state_dict=self.algo.optim_state_dict())

    def save_itr_snapshot(self, itr, sample_itr):
        self.logger.save_itr_params(itr, self.get_itr_snapshot(itr, sample_itr))

    def get_traj_info_kwargs(self):
        return dict(discount=getattr(self.algo, "discount", 1))

    def store_diagnostics(self, itr, sampler_itr, traj_infos, opt_info):
        self._traj_infos.extend(traj_infos)
        for k, v in self._opt_infos.items():
            new_v = getattr(opt_info, k, [])
            v.extend(new_v if isinstance(new_v, list) else [new_v])

    def log_diagnostics(self, itr, sampler_itr, throttle_time, prefix="Diagnostics/"):
        logger = self.logger
        self.save_itr_snapshot(itr, sampler_itr)
        new_time = time.time()
        time_elapsed = max(new_time - self._last_time, 1e-9)
        new_updates = self.algo.update_counter - self._last_update_counter
        new_samples = self.sampler.batch_size * (sampler_itr - self._last_sampler_itr)
        updates_per_second = float("nan") if itr == 0 else new_updates / time_elapsed
        samples_per_second = float("nan") if itr == 0 else new_samples / time_elapsed
        if self._eval:
            new_eval_time = self.ctrl.eval_time.value
            eval_time_elapsed = new_eval_time - self._last_eval_time
            non_eval_time_elapsed = max(time_elapsed - eval_time_elapsed, 1e-9)
            non_eval_samples_per_second = float("nan") if itr == 0 else new_samples / non_eval_time_elapsed
            self._last_eval_time = new_eval_time
        cum_steps = sampler_itr * self.sampler.batch_size       # per rank, as the reference (no * world_size)
        replay_ratio = new_updates * self.algo.batch_size * self.world_size / max(1, new_samples)
        cum_replay_ratio = self.algo.update_counter * self.algo.batch_size * self.world_size / max(1, cum_steps)
        with logger.tabular_prefix(prefix):
            logger.record_tabular("Iteration", itr)
            logger.record_tabular("SamplerIteration", sampler_itr)
            logger.record_tabular("CumTime (s)", new_time - self._start_time)
            logger.record_tabular("CumSteps", cum_steps)
            logger.record_tabular("CumUpdates", self.algo.update_counter)
            logger.record_tabular("ReplayRatio", replay_ratio)
            logger.record_tabular("CumReplayRatio", cum_replay_ratio)
            logger.record_tabular("StepsPerSecond", samples_per_second)
            if self._eval:
                logger.record_tabular("NonEvalSamplesPerSecond", non_eval_samples_per_second)
            logger.record_tabular("UpdatesPerSecond", updates_per_second)
            logger.record_tabular("OptThrottle", (time_elapsed - throttle_time) / time_elapsed)
        self._log_infos()
        self._last_time = new_time
        self._last_itr = itr
        self._last_sampler_itr = sampler_itr
        self._last_update_counter = self.algo.update_counter
        logger.dump_tabular(with_prefix=False)
        logger.log(f"Optimizing over {self.log_interval_itrs} sampler iterations.")

    def _log_infos(self, traj_infos=None):
        logger = self.logger
        if traj_infos is None:
            traj_infos = self._traj_infos
        if traj_infos:
            for k in traj_infos[0]:
                if not k.startswith("_"):
                    logger.record_tabular_misc_stat(k, [info[k] for info in traj_infos])
        if self._opt_infos:
            for k, v in self._opt_infos.items():
                logger.record_tabular_misc_stat(k, v)
        self._opt_infos = {k: list() for k in self._opt_infos}


class AsyncRl(AsyncRlBase):
    """Online performance tracking (async_rl.py:400-434)."""

    def __init__(self, *args, log_traj_window=100, **kwargs):
        super().__init__(*args, **kwargs)
        self.log_traj_window = int(log_traj_window)

    def initialize_logging(self):
        self._traj_infos = deque(maxlen=self.log_traj_window)
        self._cum_completed_trajs = 0
        self._new_completed_trajs = 0
        super().initialize_logging()

    def store_diagnostics(self, itr, sampler_itr, traj_infos, opt_info):
        self._cum_completed_trajs += len(traj_infos)
        self._new_completed_trajs += len(traj_infos)
        super().store_diagnostics(itr, sampler_itr, traj_infos, opt_info)

    def log_diagnostics(self, itr, sampler_itr, throttle_time, prefix="Diagnostics/"):
        logger = self.logger
        with logger.tabular_prefix(prefix):
            logger.record_tabular("CumCompletedTrajs", self._cum_completed_trajs)
            logger.record_tabular("NewCompletedTrajs", self._new_completed_trajs)
            logger.record_tabular("StepsInTrajWindow", sum(info["Length"] for info in self._traj_infos))
        super().log_diagnostics(itr, sampler_itr, throttle_time, prefix=prefix)
        self._new_completed_trajs = 0


class AsyncRlEval(AsyncRlBase):
    """Offline evaluation by the sampler every ``log_interval`` iterations (async_rl.py:437-461)."""

    _eval = True

    def initialize_logging(self):
        self._traj_infos = list()
        self._last_eval_time = 0.
        super().initialize_logging()

    def log_diagnostics(self, itr, sampler_itr, throttle_time, prefix="Diagnostics/"):
        logger = self.logger
        if not self._traj_infos:
            logger.log("WARNING: had no complete trajectories in eval.")
        steps_in_eval = sum(info["Length"] for info in self._traj_infos)
        with logger.tabular_prefix(prefix):
            logger.record_tabular("StepsInEval", steps_in_eval)
            logger.record_tabular("TrajsInEval", len(self._traj_infos))
            logger.record_tabular("CumEvalTime", self.ctrl.eval_time.value)
        super().log_diagnostics(itr, sampler_itr, throttle_time, prefix=prefix)
        self._traj_infos = list()


# ---------------------------------------------------------------------- worker threads (async_rl.py:512-608)
class _stream_scope:
    """Run a thread's device work on its own CUDA stream (a new thread starts on the default stream, which it would
    share with the optimizer: correct, but serialised)."""

    def __init__(self, device):
        self.device = device
        self.ctx = None

    def __enter__(self):
        if self.device is not None and torch.cuda.is_available() and torch.device(self.device).type == "cuda":
            torch.cuda.set_device(self.device)
            self.stream = torch.cuda.Stream(self.device)
            self.stream.wait_stream(torch.cuda.default_stream(self.device))   # everything set up before the threads started
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.stream.synchronize()
            self.ctx.__exit__(*exc)
        return False


def run_async_sampler(sampler, ctrl, traj_infos_queue, n_itr, device=None, errors=None, logger=None):
    """async_rl.py:512-536: toggle the double buffer each iteration; wait for the copier before writing a buffer,
    signal it after."""
    itr = -1
    try:
        with _stream_scope(device):
            db_idx = 0
            for itr in range(n_itr):
                ctrl.sample_copied[db_idx].acquire()
                if ctrl.quit.value:
                    break
                traj_infos = sampler.obtain_samples(itr, db_idx)
                ctrl.published.value += 1
                ctrl.sample_ready[db_idx].release()
                with ctrl.sampler_itr.get_lock():
                    for traj_info in traj_infos:
                        traj_infos_queue.put(traj_info)
                    ctrl.sampler_itr.value = itr
                db_idx ^= 1
        if logger is not None:
            logger.log(f"Async sampler reached final itr: {itr + 1}, quitting.")
    except BaseException as e:  # noqa: BLE001 - surfaced by the optimizer loop; never leave it waiting
        if errors is not None:
            errors.append(("sampler", e))
    finally:
        ctrl.quit.value = True                                # this ends the experiment
        try:
            sampler.shutdown()
            sampler._shut_down = True
        except Exception as e:  # noqa: BLE001
            if errors is not None:
                errors.append(("sampler shutdown", e))
        for s in ctrl.sample_ready:
            s.release()                                       # let the copier finish and quit


def run_async_sampler_eval(sampler, ctrl, traj_infos_queue, n_itr, eval_itrs, device=None, errors=None, logger=None):
    """async_rl.py:539-571."""
    itr = -1
    try:
        with _stream_scope(device):
            db_idx = 0
            for itr in range(n_itr + 1):                      # +1 for the last evaluation
                ctrl.sample_copied[db_idx].acquire()
                if ctrl.quit.value:
                    break
                sampler.obtain_samples(itr, db_idx)
                ctrl.published.value += 1
                ctrl.sample_ready[db_idx].release()
                if itr % eval_itrs == 0:
                    eval_time = -time.time()
                    traj_infos = sampler.evaluate_agent(itr)
                    eval_time += time.time()
                    ctrl.eval_time.value += eval_time
                    with ctrl.sampler_itr.get_lock():
                        for traj_info in traj_infos:
                            traj_infos_queue.put(traj_info)
                        traj_infos_queue.put(None)            # the master reads until this sentinel
                        ctrl.sampler_itr.value = itr
                else:
                    ctrl.sampler_itr.value = itr
                db_idx ^= 1
        if logger is not None:
            logger.log(f"Async sampler reached final itr: {itr + 1}, quitting.")
    except BaseException as e:  # noqa: BLE001
        if errors is not None:
            errors.append(("sampler", e))
        traj_infos_queue.put(None)                            # a master blocked on the sentinel must wake up
    finally:
        ctrl.quit.value = True
        try:
            sampler.shutdown()
            sampler._shut_down = True
        except Exception as e:  # noqa: BLE001
            if errors is not None:
                errors.append(("sampler shutdown", e))
        for s in ctrl.sample_ready:
            s.release()


def memory_copier(sampler, samples_to_buffer, replay_buffer, ctrl, device=None, errors=None, logger=None):
    """async_rl.py:574-608 - ONE copier serving both halves of the double buffer in order (the reference runs two
    processes because a shared-memory copy of a batch takes milliseconds; an HBM append takes microseconds, and a single
    copier keeps ``samples_to_buffer`` state - e.g. R2D1's input priorities - in iteration order).  Unlike the reference's
    copiers (which drop a batch that completes together with ``quit``) every published batch is appended."""
    try:
        with _stream_scope(device):
            db_idx = copied = 0
            while True:
                ctrl.sample_ready[db_idx].acquire()
                if copied >= ctrl.published.value:            # not a batch: the quit sentinel (or an error path's wake-up)
                    if ctrl.quit.value:
                        break
                    continue
                batch = sampler.acquire_batch(db_idx)         # this stream waits for the sampler's publish event
                replay_buffer.append_samples(samples_to_buffer(batch))
                sampler.release_batch(db_idx)                 # ... and the sampler's next write waits for these reads
                copied += 1
                ctrl.sample_copied[db_idx].release()
                db_idx ^= 1
        if logger is not None:
            logger.log("Memory copier shutting down.")
    except BaseException as e:  # noqa: BLE001
        if errors is not None:
            errors.append(("memory copier", e))
        ctrl.quit.value = True
        for s in ctrl.sample_copied:
            s.release()
