"""Env-side collectors (CPU): start-up decorrelation and the per-batch stepping loops that talk to
the master's step engine through the ``[B]`` step-exchange buffer.

Mirrors ``rlpyt/samplers/collectors.py:80-119`` (DecorrelatingStartCollector.start_envs),
``rlpyt/samplers/parallel/gpu/collectors.py:18-50`` (GpuResetCollector) and ``:53-126``
(GpuWaitResetCollector).  Difference by design: the collector no longer copies observations /
actions / agent_info into a host ``[T,B]`` batch (the master records them in HBM); it only writes
the next observation, reward and done into the step buffer and ``env_info`` into its host log rows.
"""
import numpy as np

from rlpyt_b200._lib import host_stream_copy, host_stream_copy_ptr, load as _load_lib


class DecorrelatingStartCollector:

    mid_batch_reset = True

    def __init__(self, rank, envs, env_info_np, batch_T, TrajInfoCls, sync=None, step_buffer_np=None,
                 global_B=1, env_ranks=None):
        self.rank, self.envs, self.env_info_np = rank, envs, env_info_np
        self.batch_T, self.TrajInfoCls = batch_T, TrajInfoCls
        self.sync, self.step_buffer_np = sync, step_buffer_np
        self.global_B, self.env_ranks = global_B, env_ranks

    def start_envs(self, max_decorrelation_steps=0):
        """reset() every env, then a random number of random actions each (collectors.py:80-119);
        leaves observation / prev_action / prev_reward in the step buffer."""
        step = self.step_buffer_np
        traj_infos = [self.TrajInfoCls() for _ in self.envs]
        for b, env in enumerate(self.envs):
            o = env.reset()
            a = env.action_space.null_value()
            r = 0
            if max_decorrelation_steps != 0:
                n_steps = 1 + int(np.random.rand() * max_decorrelation_steps)
                for _ in range(n_steps):
                    a = env.action_space.sample()
                    o, r, d, info = env.step(a)
                    traj_infos[b].step(o, a, r, d, None, info)
                    if getattr(info, "traj_done", d):
                        o = env.reset()
                        traj_infos[b] = self.TrajInfoCls()
                    if d:
                        a = env.action_space.null_value()
                        r = 0
            step.observation[b] = o
            step.action[b] = a
            step.reward[b] = r
            step.done[b] = False
        return traj_infos

    def reset_if_needed(self):
        pass


class GpuResetCollector(DecorrelatingStartCollector):
    """Resets an env immediately when its trajectory ends (gpu/collectors.py:18-50)."""

    mid_batch_reset = True

    def collect_batch(self, traj_infos, itr):
        """gpu/collectors.py:25-48.  The per-environment numpy scalar writes of the reference loop (reward, done, every
        env_info field: ~2.5 us per env step of pure indexing overhead) are gathered in Python lists and written once
        per time step - same values, same places; the observation rows are views made once per batch."""
        act_ready, obs_ready = self.sync.act_ready, self.sync.obs_ready
        step = self.step_buffer_np
        envs = self.envs
        n = len(envs)
        obs_rows = [step.observation[b] for b in range(n)]
        _load_lib()                                              # host_stream_copy_ptr calls the library directly
        row_ptrs = [row.ctypes.data for row in obs_rows] if step.observation[0].flags["C_CONTIGUOUS"] else None
        row_nbytes, row_dtype = obs_rows[0].nbytes, obs_rows[0].dtype
        info_np = self.env_info_np
        info_fields = getattr(info_np, "_fields", None)
        if info_fields is not None and not all(isinstance(getattr(info_np, f), np.ndarray) for f in info_fields):
            info_fields = None                                  # nested env_info: per-environment writes
        completed = []
        obs_ready.release()  # previous observation already in the step buffer
        for t in range(self.batch_T):
            act_ready.acquire()  # the master has written step.action
            actions = step.action.tolist() if step.action.ndim == 1 else None
            rewards, dones, infos = [None] * n, [None] * n, [None] * n
            for b in range(n):
                env = envs[b]
                a = step.action[b] if actions is None else actions[b]
                o, r, d, env_info = env.step(a)
                traj_infos[b].step(obs_rows[b], a, r, d, None, env_info)
                if getattr(env_info, "traj_done", d):
                    completed.append(traj_infos[b].terminate(o))
                    traj_infos[b] = self.TrajInfoCls()
                    o = env.reset()
                # non-temporal copy: keeps the DMA source out of this core's L2 (row address precomputed)
                if row_ptrs is None or not host_stream_copy_ptr(row_ptrs[b], row_nbytes, row_dtype, o):
                    host_stream_copy(obs_rows[b], o)
                rewards[b], dones[b], infos[b] = r, d, env_info
            step.reward[:] = rewards
            step.done[:] = dones
            if info_fields is not None and all(infos) and all(type(i) is type(infos[0]) for i in infos):
                for f, col in zip(info_fields, zip(*infos)):          # one write per field and time step
                    getattr(info_np, f)[t] = col
            else:
                for b in range(n):
                    if infos[b]:
                        info_np[t, b] = infos[b]
            obs_ready.release()
        return traj_infos, completed


class GpuWaitResetCollector(DecorrelatingStartCollector):
    """Leaves a finished env idle until the batch ends, recording blanks (gpu/collectors.py:53-126):
    ``step.done`` stays True for the rest of the batch, reward / observation are zeroed, resets
    happen between batches (``reset_if_needed``).  The master blanks the action / agent_info rows
    of those columns (the reference's worker does it on the host batch, :86-92, :107-111).
    Quirk kept on purpose: at the start of a batch the held terminal observation is reinstated for
    every column whose ``done`` flag is still set (:73-75), even if ``reset_if_needed`` just wrote a
    fresh reset observation there."""

    mid_batch_reset = False

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.need_reset = np.zeros(len(self.envs), dtype=bool)
        self.temp_observation = self.step_buffer_np.observation.copy()

    def collect_batch(self, traj_infos, itr):
        act_ready, obs_ready = self.sync.act_ready, self.sync.obs_ready
        step = self.step_buffer_np
        completed = []
        held = np.where(step.done)[0]
        step.observation[held] = self.temp_observation[held]
        step.done[:] = False  # resets were done between batches
        obs_ready.release()
        for t in range(self.batch_T):
            act_ready.acquire()
            for b, env in enumerate(self.envs):
                if step.done[b]:
                    step.action[b] = 0  # record blank; step.done[b] stays True
                    step.reward[b] = 0
                    continue
                o, r, d, env_info = env.step(step.action[b])
                traj_infos[b].step(step.observation[b], step.action[b], r, d, None, env_info)
                if getattr(env_info, "traj_done", d):
                    completed.append(traj_infos[b].terminate(o))
                    traj_infos[b] = self.TrajInfoCls()
                    self.need_reset[b] = True
                if d:
                    self.temp_observation[b] = o  # held until the next batch starts
                    o = 0
                step.observation[b] = o
                step.reward[b] = r
                step.done[b] = d
                if env_info:
                    self.env_info_np[t, b] = env_info
            obs_ready.release()
        return traj_infos, completed

    def reset_if_needed(self):
        step = self.step_buffer_np
        for b in np.where(self.need_reset)[0]:
            step.observation[b] = self.envs[b].reset()
            step.action[b] = 0
            step.reward[b] = 0
        self.need_reset[:] = False
