"""DQN / Double-DQN with n-step returns and prioritized frame replay on device-resident data
(mirror of ``rlpyt/algos/dqn/dqn.py:19-283``: same constructor, ``initialize``, ``optimize_agent``,
``samples_to_buffer``, ``loss``).  SURVEY.md section 8(f) row 1: this turns the replay path
(sum-tree + frame gather kernels) into an end-to-end learner.

What changes against the reference, behind the same methods:
* the replay buffer, its fp64 sum-tree and the sampled batch live in HBM (rlpyt_b200.replays), so
  ``sample_batch`` is a few kernel launches instead of a 15 ms Python loop;
* both networks' outputs stay on the device and the loss arithmetic + its gradient + the new
  priorities are one fused kernel (csrc/dqn_loss.cu) instead of ~20 torch-CPU ops after two D2H copies;
* ``td_abs_errors`` go straight into ``update_batch_priorities`` (device pow + tree update);
* clip + Adam is ``FlatAdam.clip_and_step``; OptInfo rows are read back once per ``optimize_agent``.
"""
from collections import namedtuple

import numpy as np
import torch

from rlpyt_b200.algos.base import RlAlgorithm
from rlpyt_b200.algos.dqn import loss_ops
from rlpyt_b200.algos.optim import FlatAdam
from rlpyt_b200.replays.non_sequence.frame import (AsyncPrioritizedReplayFrameBuffer, AsyncUniformReplayFrameBuffer,
                                                   PrioritizedReplayFrameBuffer, UniformReplayFrameBuffer)
from rlpyt_b200.utils.collections import namedarraytuple

OptInfo = namedtuple("OptInfo", ["loss", "gradNorm", "tdAbsErr"])
SamplesToBuffer = namedarraytuple("SamplesToBuffer", ["observation", "action", "reward", "done"])


class DQN(RlAlgorithm):

    opt_info_fields = tuple(f for f in OptInfo._fields)

    def __init__(self, discount=0.99, batch_size=32, min_steps_learn=int(5e4), delta_clip=1., replay_size=int(1e6),
                 replay_ratio=8, target_update_tau=1, target_update_interval=312, n_step_return=1,
                 learning_rate=2.5e-4, OptimCls=FlatAdam, optim_kwargs=None, initial_optim_state_dict=None,
                 clip_grad_norm=10., eps_steps=int(1e6), double_dqn=False, prioritized_replay=False, pri_alpha=0.6,
                 pri_beta_init=0.4, pri_beta_final=1., pri_beta_steps=int(50e6), default_priority=None,
                 ReplayBufferCls=None, updates_per_sync=1):
        if optim_kwargs is None:
            optim_kwargs = dict(eps=0.01 / batch_size)                     # dqn.py:66-67
        if default_priority is None:
            default_priority = delta_clip                                  # dqn.py:68-69
        self._batch_size = batch_size
        self.discount, self.min_steps_learn, self.delta_clip = discount, min_steps_learn, delta_clip
        self.replay_size, self.replay_ratio = replay_size, replay_ratio
        self.target_update_tau, self.target_update_interval = target_update_tau, target_update_interval
        self.n_step_return, self.learning_rate = n_step_return, learning_rate
        self.OptimCls, self.optim_kwargs = OptimCls, optim_kwargs
        self.initial_optim_state_dict, self.clip_grad_norm = initial_optim_state_dict, clip_grad_norm
        self.eps_steps, self.double_dqn, self.prioritized_replay = eps_steps, double_dqn, prioritized_replay
        self.pri_alpha, self.pri_beta_init, self.pri_beta_final = pri_alpha, pri_beta_init, pri_beta_final
        self.pri_beta_steps, self.default_priority = pri_beta_steps, default_priority
        self.ReplayBufferCls, self.updates_per_sync = ReplayBufferCls, updates_per_sync
        self.update_counter = 0

    def initialize(self, agent, n_itr, batch_spec, mid_batch_reset, examples, world_size=1, rank=0):
        """dqn.py:74-93."""
        self.agent = agent
        self.n_itr = n_itr
        self.sampler_bs = sampler_bs = batch_spec.size
        self.mid_batch_reset = mid_batch_reset
        self.world_size = world_size
        self.updates_per_optimize = max(1, round(self.replay_ratio * sampler_bs / self.batch_size))
        self.min_itr_learn = int(self.min_steps_learn // sampler_bs)
        eps_itr_max = max(1, int(self.eps_steps // sampler_bs))
        agent.set_epsilon_itr_min_max(self.min_itr_learn, eps_itr_max)
        self.initialize_replay_buffer(examples, batch_spec)
        self.optim_initialize(rank)

    def async_initialize(self, agent, sampler_n_itr, batch_spec, mid_batch_reset, examples, world_size=1):
        """dqn.py:95-109: used by the asynchronous runner only; allocates the replay buffer (its lock-and-fence
        guarded variant, in HBM), leaves the optimizer to ``optim_initialize``; returns the buffer."""
        self.agent = agent
        self.n_itr = sampler_n_itr
        self.initialize_replay_buffer(examples, batch_spec, async_=True)
        self.mid_batch_reset = mid_batch_reset
        self.sampler_bs = sampler_bs = batch_spec.size
        self.world_size = world_size
        self.updates_per_optimize = self.updates_per_sync
        self.min_itr_learn = int(self.min_steps_learn // sampler_bs)
        eps_itr_max = max(1, int(self.eps_steps // sampler_bs))
        agent.set_epsilon_itr_min_max(self.min_itr_learn, eps_itr_max)
        return self.replay_buffer

    def optim_initialize(self, rank=0):
        """dqn.py:111-120."""
        self.rank = rank
        self.optimizer = self.OptimCls(self.agent.parameters(), lr=self.learning_rate, **self.optim_kwargs)
        if isinstance(self.optimizer, FlatAdam):
            self.optimizer.set_world_size(self.world_size)
        if self.initial_optim_state_dict is not None:
            self.optimizer.load_state_dict(self.initial_optim_state_dict)
        if self.prioritized_replay:
            self.pri_beta_itr = max(1, self.pri_beta_steps // self.sampler_bs)

    def initialize_replay_buffer(self, examples, batch_spec, async_=False):
        """dqn.py:122-156."""
        replay_kwargs = dict(example=self.examples_to_buffer(examples), size=self.replay_size, B=batch_spec.B,
                             discount=self.discount, n_step_return=self.n_step_return)
        if self.prioritized_replay:
            replay_kwargs.update(alpha=self.pri_alpha, beta=self.pri_beta_init, default_priority=self.default_priority)
            ReplayCls = AsyncPrioritizedReplayFrameBuffer if async_ else PrioritizedReplayFrameBuffer
        else:
            ReplayCls = AsyncUniformReplayFrameBuffer if async_ else UniformReplayFrameBuffer
        if self.ReplayBufferCls is not None:
            ReplayCls = self.ReplayBufferCls
        dev = getattr(self.agent, "device", None)
        if dev is not None and torch.device(dev).type == "cuda":
            replay_kwargs["device"] = dev
        self.replay_buffer = ReplayCls(**replay_kwargs)

    def optimize_agent(self, itr, samples=None, sampler_itr=None):
        """dqn.py:158-190."""
        itr = itr if sampler_itr is None else sampler_itr
        if samples is not None:
            self.replay_buffer.append_samples(self.samples_to_buffer(samples))
        opt_info = OptInfo(*([] for _ in range(len(OptInfo._fields))))
        if itr < self.min_itr_learn:
            return opt_info
        fused_opt = isinstance(self.optimizer, FlatAdam)
        rows = []
        for _ in range(self.updates_per_optimize):
            samples_from_replay = self.replay_buffer.sample_batch(self.batch_size)
            self.optimizer.zero_grad()
            loss, td_abs_errors = self.loss(samples_from_replay)
            loss.backward()
            if fused_opt:
                grad_norm = self.optimizer.clip_and_step(self.clip_grad_norm)
            else:
                grad_norm = torch.nn.utils.clip_grad_norm_(self.agent.parameters(), self.clip_grad_norm)
                self.optimizer.step()
            if self.prioritized_replay:
                self.replay_buffer.update_batch_priorities(td_abs_errors)
            rows.append(torch.cat([loss.detach().reshape(1), grad_norm.detach().reshape(1).to(loss.dtype),
                                   td_abs_errors.detach().reshape(-1)[::8]]))      # dqn.py:185 downsample
            self.update_counter += 1
            if self.update_counter % self.target_update_interval == 0:
                self.agent.update_target(self.target_update_tau)
        host = torch.stack(rows).cpu().numpy().astype(np.float64)                   # the call's single D2H sync
        opt_info.loss.extend(host[:, 0].tolist())
        opt_info.gradNorm.extend(host[:, 1].tolist())
        opt_info.tdAbsErr.extend(host[:, 2:].astype(np.float32).reshape(-1))
        self.update_itr_hyperparams(itr)
        return opt_info

    def examples_to_buffer(self, examples):
        return SamplesToBuffer(observation=examples["observation"], action=examples["action"],
                               reward=examples["reward"], done=examples["done"])

    def samples_to_buffer(self, samples):
        """dqn.py:200-209."""
        return SamplesToBuffer(observation=samples.env.observation, action=samples.agent.action,
                               reward=samples.env.reward, done=samples.env.done)

    def loss(self, samples):
        """dqn.py:211-265 -> (loss, td_abs_errors), both on the device."""
        if not self.mid_batch_reset:
            raise NotImplementedError                                         # as the reference, dqn.py:255-258
        qs = self.agent(*samples.agent_inputs)
        with torch.no_grad():
            target_qs = self.agent.target(*samples.target_inputs)
            next_qs = self.agent(*samples.target_inputs) if self.double_dqn else None
        is_weights = samples.is_weights if self.prioritized_replay else None
        return loss_ops.dqn_loss(qs, target_qs, next_qs, samples.action, samples.return_, samples.done_n,
                                 is_weights, self.discount ** self.n_step_return, self.delta_clip)

    def update_itr_hyperparams(self, itr):
        """dqn.py:267-283: anneal the importance-sampling exponent."""
        if self.prioritized_replay and itr <= self.pri_beta_itr:
            prog = min(1, max(0, itr - self.min_itr_learn) / (self.pri_beta_itr - self.min_itr_learn))
            new_beta = prog * self.pri_beta_final + (1 - prog) * self.pri_beta_init
            self.replay_buffer.set_beta(new_beta)
