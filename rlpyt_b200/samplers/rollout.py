"""Master-side step engine shared by the serial and the parallel GPU sampler: moves one step of
observations pinned-host -> HBM, runs ``agent.step`` on the resident slices, records the
outputs in the ``[T,B]`` device buffers and hands the actions back to the env side.

This is the device half of ``ActionServer.serve_actions`` (rlpyt/samplers/parallel/gpu/
action_server.py:17-74) and of ``CpuResetCollector.collect_batch`` (rlpyt/samplers/parallel/cpu/
collectors.py:25-65); the env half stays on the CPU.
"""
import os

import numpy as np
import torch


class DeviceRollout:

    def __init__(self, samples, host, agent, device, stream=None):
        self.samples, self.host, self.agent = samples, host, agent
        self.device = device
        self.side_stream = stream       # alternating sampler: one stream per half, so one half's H2D overlaps the other's compute
        self.step_np, self.step_pyt = host["step_np"], host["step_pyt"]
        self.all_action, self.all_reward = host["all_action"], host["all_reward"]
        B = self.step_np.action.shape[0]
        self.B = B
        self.T = samples.env.done.shape[0]
        # agent inputs of the current step (prev_action / prev_reward), resident
        self.in_action = torch.zeros_like(self.all_action[0])
        self.in_reward = torch.zeros_like(self.all_reward[0])
        self.obs_extra = torch.zeros_like(samples.env.observation[0])  # final obs (bootstrap)
        self.done_step = torch.zeros(B, dtype=torch.bool, device=device)
        self.stream = torch.cuda.current_stream(device)
        # One CUDA graph per time step t (H2D of the step buffer -> agent.step -> record row t ->
        # D2H of the actions): ~30 small launches become one graph launch per env step.  Captured
        # lazily after an eager warm-up batch (cuDNN autotuning cannot run under capture).
        self.use_graphs = os.environ.get("RLPYT_B200_SAMPLER_GRAPHS", "1") == "1"
        self._warned_unpinned = False
        self.capture_error_mode = "global"
        # RLPYT_B200_SAMPLER_ZEROCOPY=1: no H2D of the observations in front of agent.step - the first layer reads the
        # frames out of the page-locked step buffer itself and records them in observation[t] on the way (models that
        # set ``accepts_host_mapped_frames``; feed-forward agents; the buffer must be page-locked)
        self.zero_copy = (os.environ.get("RLPYT_B200_SAMPLER_ZEROCOPY", "0") == "1" and bool(host.get("pinned", False))
                          and bool(getattr(getattr(agent, "model", None), "accepts_host_mapped_frames", False))
                          and not getattr(agent, "recurrent", False) and self.step_np.observation.dtype == np.uint8
                          and self.step_np.observation.ndim == 4)
        self._host_obs_ptr = int(self.step_np.observation.ctypes.data)
        self._act_event = None
        self._graphs = {}
        self._eager_batches = 0

    # ---- H2D of what the envs produced since the last step ---------------------------------------
    def obs_slot(self, k):
        """Where observation(k) lives in HBM: row k of the batch, or the bootstrap slot for k == T."""
        return self.samples.env.observation[k] if k < self.T else self.obs_extra

    def set_worker_chunks(self, slices):
        """``slices``: each env worker's rows of THIS engine's B range.  Enables ``upload_worker_rows``: the observation
        rows of one worker go to HBM as soon as that worker has signalled (one thin cudaMemcpyAsync through the C ABI),
        so the H2D of a step overlaps the workers that are still stepping and only the last worker's rows are on the
        critical path; ``upload_async(k, ..., obs_done=True)`` then moves the small fields only."""
        from rlpyt_b200 import _lib
        obs = self.step_np.observation
        self._row_bytes = int(obs[0].nbytes)
        self._chunks = [(int(sl.start), int(sl.stop - sl.start)) for sl in slices]
        self._src_base = int(obs.ctypes.data)
        self._upload_fn = _lib.load().rl_upload_async
        self._dst_base = {}

    def upload_worker_rows(self, k, i):
        start, n = self._chunks[i]
        base = self._dst_base.get(k)
        if base is None:
            base = self._dst_base[k] = int(self.obs_slot(k).data_ptr())
        off = start * self._row_bytes
        stream = self.side_stream if self.side_stream is not None else torch.cuda.current_stream(self.device)
        rc = self._upload_fn(base + off, self._src_base + off, n * self._row_bytes, stream.cuda_stream)
        if rc != 0:
            from rlpyt_b200 import _lib
            _lib.check(rc, "rl_upload_async")

    def upload(self, k, zero_inputs_on_done, obs_done=False):
        """Event k (0..T): the envs have written observation(k), reward(k-1), done(k-1) into the
        step buffer.  Record them at their [T,B] rows and stage the agent inputs."""
        s = self.samples
        obs_dst = self.obs_slot(k)
        if not obs_done and not (self.zero_copy and k < self.T):   # zero-copy: agent.step's first layer brings the frames in
            obs_dst.copy_(self.step_pyt.observation, non_blocking=True)
        self.all_reward[k].copy_(self.step_pyt.reward, non_blocking=True)       # reward(k-1) = prev_reward(k)
        self.done_step.copy_(self.step_pyt.done, non_blocking=True)
        if k >= 1:
            s.env.done[k - 1].copy_(self.done_step, non_blocking=True)
        self.in_reward.copy_(self.all_reward[k], non_blocking=True)
        if zero_inputs_on_done:                                                   # action_server.py:49-53
            self.in_action.masked_fill_(self.done_step, 0)
            self.in_reward.masked_fill_(self.done_step, 0)
        return obs_dst

    def begin_batch(self):
        """Row 0 of prev_action is what the agent saw as previous action when the batch starts
        (gpu/collectors.py:23-24)."""
        self.all_action[0].copy_(self.in_action, non_blocking=True)

    # ---- agent.step on resident data ---------------------------------------------------------------
    @torch.no_grad()
    def act(self, t, obs_dev, blank_done_rows=False, sync=True):
        if self.zero_copy:
            from rlpyt_b200.utils.gather import HostMappedFrames
            obs_dev = HostMappedFrames(self._host_obs_ptr, self.obs_slot(t))
        step = self.agent.step(obs_dev, self.in_action, self.in_reward)
        action, agent_info = step.action, step.agent_info
        if blank_done_rows:  # wait-reset collectors record blanks for finished envs
            keep = ~self.done_step
            action = action * keep
            agent_info = _mask_rows(agent_info, keep)
        self.samples.agent.action[t].copy_(action, non_blocking=True)
        self.samples.agent.agent_info[t] = agent_info
        self.in_action.copy_(action, non_blocking=True)
        self.step_pyt.action.copy_(action, non_blocking=True)                    # D2H for the envs
        if sync:
            torch.cuda.current_stream(self.device).synchronize()                  # actions are on the host

    # ---- one env step = upload + agent.step + record (+ D2H), optionally replayed as a CUDA graph ----
    def _step_body(self, k, zero_inputs_on_done, blank_done_rows, obs_done=False):
        obs_dev = self.upload(k, zero_inputs_on_done, obs_done=obs_done)
        if k == 0:
            self.begin_batch()
        self.act(k, obs_dev, blank_done_rows=blank_done_rows, sync=False)

    def _bootstrap_body(self):
        obs_dev = self.upload(self.T, zero_inputs_on_done=False)
        self.bootstrap(obs_dev)

    def _run(self, key, body):
        if self.side_stream is not None:
            with torch.cuda.stream(self.side_stream):
                return self._run_on_current(key, body)
        return self._run_on_current(key, body)

    def _run_on_current(self, key, body):
        if not (self.use_graphs and self._eager_batches >= 1 and getattr(self.agent, "device", None) is not None
                and self.agent.device.type == "cuda") or getattr(self.agent, "recurrent", False):
            # recurrent agents carry their state in fresh tensors from step to step (and zero columns of it on
            # episode ends): not a fixed-address pattern a captured graph could replay
            body()
            return
        if not self.host.get("pinned", False):
            # the step body copies out of / into the host step buffer with non_blocking=True: only page-locked
            # memory may be captured (pin_shared() can fail, e.g. under a locked-memory ulimit) -> stay eager
            if not self._warned_unpinned:
                import warnings
                warnings.warn("rlpyt_b200: the step buffer is not page-locked; sampler steps run eagerly (no CUDA graphs)")
                self._warned_unpinned = True
            body()
            return
        g = self._graphs.get(key)
        if g is None:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            issuing = torch.cuda.current_stream(self.device)
            try:
                # "thread_local": in the asynchronous runner other threads of this process keep issuing CUDA calls
                # (synchronisations, allocations) on their own streams while this thread captures
                with torch.cuda.graph(g, stream=self.side_stream, capture_error_mode=self.capture_error_mode):
                    body()
            except RuntimeError as e:
                # A capture can be invalidated from outside the body (another thread's or a finalizer's CUDA call in
                # "global" mode, a library that allocates on first use): nothing of the body has executed on the
                # device, so the step is simply issued eagerly - and stays eager for this engine.
                import warnings
                warnings.warn(f"rlpyt_b200: CUDA-graph capture of the sampler step failed ({str(e).splitlines()[0][:120]}); "
                              "this sampler steps eagerly from now on")
                self.use_graphs = False
                self._graphs.clear()
                torch.cuda.set_stream(issuing)               # torch.cuda.graph.__exit__ raised before it restored the stream
                torch.cuda.synchronize(self.device)
                body()
                return
            self._graphs[key] = g
        g.replay()

    # ---- the same step in two launches (alternating sampler): upload_async(k) may be issued while the OTHER half's
    # act is still running on its own stream; act_async(k) follows on this half's stream; wait() = actions on the host
    def upload_async(self, k, zero_inputs_on_done, obs_done=False):
        def body():
            self.upload(k, zero_inputs_on_done, obs_done=obs_done)
            if k == 0:
                self.begin_batch()
        self._run(("up", k, zero_inputs_on_done, obs_done), body)

    def act_async(self, k, blank_done_rows=False):
        self._run(("act", k, blank_done_rows), lambda: self.act(k, self.obs_slot(k), blank_done_rows=blank_done_rows, sync=False))
        if self.side_stream is not None:
            if self._act_event is None:
                self._act_event = torch.cuda.Event()
            self._act_event.record(self.side_stream)

    def act_done(self):
        """Non-blocking: have the actions of the last ``act_async`` reached the step buffer?"""
        return self._act_event is None or self._act_event.query()

    def wait(self):
        (self.side_stream or torch.cuda.current_stream(self.device)).synchronize()

    def step(self, k, zero_inputs_on_done, blank_done_rows=False, obs_done=False):
        """Event k in [0,T): observation(k) is in the step buffer (``obs_done``: already uploaded per worker by
        ``upload_worker_rows``) -> actions are on the host on return."""
        self._run((k, zero_inputs_on_done, blank_done_rows, obs_done),
                  lambda: self._step_body(k, zero_inputs_on_done, blank_done_rows, obs_done))
        self.wait()

    def finish(self):
        """Event T: final observation -> bootstrap value."""
        self._run(("bootstrap",), self._bootstrap_body)

    def end_batch(self):
        self._eager_batches += 1

    def zero_inputs_where_done(self):
        self.in_action.masked_fill_(self.done_step, 0)
        self.in_reward.masked_fill_(self.done_step, 0)

    @torch.no_grad()
    def bootstrap(self, obs_dev):
        if "bootstrap_value" in self.samples.agent:
            self.samples.agent.bootstrap_value[0].copy_(
                self.agent.value(obs_dev, self.in_action, self.in_reward), non_blocking=True)


def _mask_rows(buf, keep):
    """Zero the rows (leading dim B) of every tensor of a namedarraytuple where ``keep`` is False."""
    if isinstance(buf, torch.Tensor):
        k = keep.view((-1,) + (1,) * (buf.dim() - 1))
        return buf * k.to(buf.dtype)
    return buf._make(tuple(_mask_rows(b, keep) for b in buf))
