"""R2D1: recurrent-replay DQN (mirror of ``rlpyt/algos/dqn/r2d1.py:22-345``: same constructor, ``initialize_replay_buffer``,
``optimize_agent``, ``samples_to_buffer``, ``compute_input_priorities``, ``loss``, ``value_scale`` / ``inv_value_scale``).
SURVEY.md section 8(f) row 4.

Behind the same methods, against the reference:
* the sequence replay (frames, stored RNN states, fp64 sum-tree) lives in HBM (rlpyt_b200.replays.sequence): a batch of
  ``batch_B`` sequences of ``warmup_T + batch_T + n_step`` frame stacks is two kernel launches, not a Python loop per sample;
* the sampled batch, both networks' outputs and every step of the loss stay on the device (the reference copies the
  Q-values of three forward passes to the CPU and does the TD arithmetic there, r2d1.py:283-322);
* input priorities of incoming samples (r2d1.py:171-227) are computed on the device from the recorded Q-values;
* clip + Adam is ``FlatAdam.clip_and_step``; OptInfo rows are read back once per ``optimize_agent``.
"""
from collections import namedtuple

import numpy as np
import torch

from rlpyt_b200.agents.base import AgentInputs
from rlpyt_b200.algos.dqn.dqn import DQN, SamplesToBuffer
from rlpyt_b200.algos.optim import FlatAdam
from rlpyt_b200.algos.utils import discount_return_n_step, valid_from_done
from rlpyt_b200.replays.sequence.frame import (AsyncPrioritizedSequenceReplayFrameBuffer,
                                               AsyncUniformSequenceReplayFrameBuffer,
                                               PrioritizedSequenceReplayFrameBuffer, UniformSequenceReplayFrameBuffer)
from rlpyt_b200.utils.collections import namedarraytuple
from rlpyt_b200.utils.tensor import select_at_indexes, valid_mean

OptInfo = namedtuple("OptInfo", ["loss", "gradNorm", "tdAbsErr", "priority"])
SamplesToBufferRnn = namedarraytuple("SamplesToBufferRnn", SamplesToBuffer._fields + ("prev_rnn_state",))
PrioritiesSamplesToBuffer = namedarraytuple("PrioritiesSamplesToBuffer", ["priorities", "samples"])


def _map(state, fn):
    return None if state is None else type(state)(*(fn(x) for x in state))


class R2D1(DQN):

    opt_info_fields = tuple(f for f in OptInfo._fields)

    def __init__(self, discount=0.997, batch_T=80, batch_B=64, warmup_T=40, store_rnn_state_interval=40,
                 min_steps_learn=int(1e5), delta_clip=None, replay_size=int(1e6), replay_ratio=1,
                 target_update_interval=2500, n_step_return=5, learning_rate=1e-4, OptimCls=FlatAdam, optim_kwargs=None,
                 initial_optim_state_dict=None, clip_grad_norm=80., eps_steps=int(1e6), double_dqn=True,
                 prioritized_replay=True, pri_alpha=0.6, pri_beta_init=0.9, pri_beta_final=0.9, pri_beta_steps=int(50e6),
                 pri_eta=0.9, default_priority=None, input_priorities=True, input_priority_shift=None,
                 value_scale_eps=1e-3, ReplayBufferCls=None, updates_per_sync=1):
        if optim_kwargs is None:
            optim_kwargs = dict(eps=1e-3)                                  # r2d1.py:79-80
        if default_priority is None:
            default_priority = delta_clip or 1.                            # :81-82
        if input_priority_shift is None:
            input_priority_shift = warmup_T // store_rnn_state_interval    # :83-84
        self.discount, self.batch_T, self.batch_B, self.warmup_T = discount, batch_T, batch_B, warmup_T
        self.store_rnn_state_interval, self.min_steps_learn, self.delta_clip = store_rnn_state_interval, min_steps_learn, delta_clip
        self.replay_size, self.replay_ratio = replay_size, replay_ratio
        self.target_update_interval, self.n_step_return, self.learning_rate = target_update_interval, n_step_return, learning_rate
        self.target_update_tau = 1
        self.OptimCls, self.optim_kwargs, self.initial_optim_state_dict = OptimCls, optim_kwargs, initial_optim_state_dict
        self.clip_grad_norm, self.eps_steps, self.double_dqn = clip_grad_norm, eps_steps, double_dqn
        self.prioritized_replay, self.pri_alpha = prioritized_replay, pri_alpha
        self.pri_beta_init, self.pri_beta_final, self.pri_beta_steps = pri_beta_init, pri_beta_final, pri_beta_steps
        self.pri_eta, self.default_priority = pri_eta, default_priority
        self.input_priorities, self.input_priority_shift = input_priorities, input_priority_shift
        self.value_scale_eps, self.ReplayBufferCls, self.updates_per_sync = value_scale_eps, ReplayBufferCls, updates_per_sync
        self._batch_size = (self.batch_T + self.warmup_T) * self.batch_B   # :86
        self.update_counter = 0

    def initialize_replay_buffer(self, examples, batch_spec, async_=False):
        """r2d1.py:88-134."""
        example_to_buffer = SamplesToBuffer(observation=examples["observation"], action=examples["action"],
                                            reward=examples["reward"], done=examples["done"])
        if self.store_rnn_state_interval > 0:
            example_to_buffer = SamplesToBufferRnn(*example_to_buffer, prev_rnn_state=examples["agent_info"].prev_rnn_state)
        replay_kwargs = dict(example=example_to_buffer, size=self.replay_size, B=batch_spec.B, discount=self.discount,
                             n_step_return=self.n_step_return, rnn_state_interval=self.store_rnn_state_interval,
                             batch_T=self.batch_T + self.warmup_T)
        if self.prioritized_replay:
            replay_kwargs.update(alpha=self.pri_alpha, beta=self.pri_beta_init, default_priority=self.default_priority,
                                 input_priorities=self.input_priorities, input_priority_shift=self.input_priority_shift)
            ReplayCls = AsyncPrioritizedSequenceReplayFrameBuffer if async_ else PrioritizedSequenceReplayFrameBuffer
        else:
            ReplayCls = AsyncUniformSequenceReplayFrameBuffer if async_ else UniformSequenceReplayFrameBuffer
        if self.ReplayBufferCls is not None:
            ReplayCls = self.ReplayBufferCls
        dev = getattr(self.agent, "device", None)
        if dev is not None and torch.device(dev).type == "cuda":
            replay_kwargs["device"] = dev
        self.replay_buffer = ReplayCls(**replay_kwargs)
        return self.replay_buffer

    def optimize_agent(self, itr, samples=None, sampler_itr=None):
        """r2d1.py:136-169."""
        itr = itr if sampler_itr is None else sampler_itr
        if samples is not None:
            self.replay_buffer.append_samples(self.samples_to_buffer(samples))
        opt_info = OptInfo(*([] for _ in range(len(OptInfo._fields))))
        if itr < self.min_itr_learn:
            return opt_info
        fused_opt = isinstance(self.optimizer, FlatAdam)
        rows = []
        for _ in range(self.updates_per_optimize):
            samples_from_replay = self.replay_buffer.sample_batch(self.batch_B)
            self.optimizer.zero_grad()
            loss, td_abs_errors, priorities = self.loss(samples_from_replay)
            loss.backward()
            if fused_opt:
                grad_norm = self.optimizer.clip_and_step(self.clip_grad_norm)
            else:
                grad_norm = torch.nn.utils.clip_grad_norm_(self.agent.parameters(), self.clip_grad_norm)
                self.optimizer.step()
            if self.prioritized_replay:
                self.replay_buffer.update_batch_priorities(priorities)
            rows.append(torch.cat([loss.detach().reshape(1), torch.as_tensor(grad_norm, device=loss.device).detach().reshape(1).to(loss.dtype),
                                   priorities.detach().reshape(-1), td_abs_errors.detach()[::8].reshape(-1)]))   # :163 downsample
            self.update_counter += 1
            if self.update_counter % self.target_update_interval == 0:
                self.agent.update_target()
        host = torch.stack(rows).cpu().numpy().astype(np.float64)            # the call's single D2H sync
        nB = self.batch_B
        opt_info.loss.extend(host[:, 0].tolist())
        opt_info.gradNorm.extend(host[:, 1].tolist())
        opt_info.priority.extend(host[:, 2:2 + nB].astype(np.float32).reshape(-1))
        opt_info.tdAbsErr.extend(host[:, 2 + nB:].astype(np.float32).reshape(-1))
        self.update_itr_hyperparams(itr)
        return opt_info

    def samples_to_buffer(self, samples):
        """r2d1.py:171-180."""
        samples_to_buffer = super().samples_to_buffer(samples)
        if self.store_rnn_state_interval > 0:
            samples_to_buffer = SamplesToBufferRnn(*samples_to_buffer, prev_rnn_state=samples.agent.agent_info.prev_rnn_state)
        if self.input_priorities:
            samples_to_buffer = PrioritiesSamplesToBuffer(priorities=self.compute_input_priorities(samples),
                                                          samples=samples_to_buffer)
        return samples_to_buffer

    @torch.no_grad()
    def compute_input_priorities(self, samples):
        """r2d1.py:182-227: n-step TD errors of the incoming [T,B] samples from the Q-values the sampler recorded ->
        one priority per column, eta * max + (1 - eta) * mean over the valid steps.  Device tensor [B]."""
        dev = self.replay_buffer.device if hasattr(self, "replay_buffer") else samples.agent.agent_info.q.device
        q = torch.as_tensor(samples.agent.agent_info.q).to(dev)
        action = torch.as_tensor(samples.agent.action).to(dev)
        reward = torch.as_tensor(samples.env.reward).to(dev)
        done = torch.as_tensor(samples.env.done).to(dev)
        q_max = torch.max(q, dim=-1).values
        q_at_a = select_at_indexes(action, q)
        return_n, done_n = discount_return_n_step(reward=reward, done=done, n_step=self.n_step_return, discount=self.discount,
                                                  do_truncated=False)
        nm1 = max(1, self.n_step_return - 1)                                  # :216
        y = self.value_scale(return_n + (1 - done_n.float()) * self.inv_value_scale(q_max[nm1:]))
        delta = abs(q_at_a[:-nm1] - y)
        if self.delta_clip is not None:
            delta = torch.clamp(delta, 0, self.delta_clip)
        valid = valid_from_done(done[:-nm1].contiguous())
        max_d = torch.max(delta * valid, dim=0).values
        mean_d = valid_mean(delta, valid, dim=0)
        return self.pri_eta * max_d + (1 - self.pri_eta) * mean_d

    def loss(self, samples):
        """r2d1.py:229-330 -> (loss, valid td_abs_errors [T,B], priorities [B]), all on the device."""
        dev = self.agent.device
        all_observation, all_action, all_reward = (x.to(dev) for x in (samples.all_observation, samples.all_action,
                                                                       samples.all_reward))
        wT, bT = self.warmup_T, self.batch_T
        if wT > 0:
            warmup_inputs = AgentInputs(observation=all_observation[:wT], prev_action=all_action[:wT],
                                        prev_reward=all_reward[:wT])
        agent_slice, target_slice = slice(wT, wT + bT), slice(wT, None)
        agent_inputs = AgentInputs(observation=all_observation[agent_slice], prev_action=all_action[agent_slice],
                                   prev_reward=all_reward[agent_slice])
        target_inputs = AgentInputs(observation=all_observation[target_slice], prev_action=all_action[target_slice],
                                    prev_reward=all_reward[target_slice])
        action = all_action[wT + 1:wT + 1 + bT]
        return_ = samples.return_[wT:wT + bT].to(dev)
        done_n = samples.done_n[wT:wT + bT].to(dev)
        done = samples.done.to(dev)
        if self.store_rnn_state_interval == 0:
            init_rnn_state = None
        else:                                                                 # [B,N,H] -> [N,B,H]
            init_rnn_state = _map(samples.init_rnn_state, lambda x: x.to(dev).transpose(0, 1).contiguous())
        if wT > 0:
            with torch.no_grad():
                _, target_rnn_state = self.agent.target(*warmup_inputs, init_rnn_state)
                _, init_rnn_state = self.agent(*warmup_inputs, init_rnn_state)
            warmup_invalid_mask = valid_from_done(done[:wT].contiguous())[-1] == 0          # [B]
            keep = (~warmup_invalid_mask).to(torch.float32)[None, :, None]
            init_rnn_state = _map(init_rnn_state, lambda x: x * keep)
            target_rnn_state = _map(target_rnn_state, lambda x: x * keep)
        else:
            target_rnn_state = init_rnn_state

        qs, _ = self.agent(*agent_inputs, init_rnn_state)                     # [T,B,A]
        q = select_at_indexes(action, qs)
        with torch.no_grad():
            target_qs, _ = self.agent.target(*target_inputs, target_rnn_state)
            if self.double_dqn:
                next_qs, _ = self.agent(*target_inputs, init_rnn_state)
                next_a = torch.argmax(next_qs, dim=-1)
                target_q = select_at_indexes(next_a, target_qs)
            else:
                target_q = torch.max(target_qs, dim=-1).values
            target_q = target_q[-bT:]

        disc = self.discount ** self.n_step_return
        y = self.value_scale(return_ + (1 - done_n.float()) * disc * self.inv_value_scale(target_q))
        delta = y - q
        losses = 0.5 * delta ** 2
        abs_delta = abs(delta)
        if self.delta_clip is not None:
            b = self.delta_clip * (abs_delta - self.delta_clip / 2)
            losses = torch.where(abs_delta <= self.delta_clip, losses, b)
        if self.prioritized_replay:
            losses = losses * samples.is_weights.to(dev).unsqueeze(0)
        valid = valid_from_done(done[wT:].contiguous())
        loss = valid_mean(losses, valid)
        td_abs_errors = abs_delta.detach()
        if self.delta_clip is not None:
            td_abs_errors = torch.clamp(td_abs_errors, 0, self.delta_clip)
        valid_td_abs_errors = td_abs_errors * valid
        max_d = torch.max(valid_td_abs_errors, dim=0).values
        mean_d = valid_mean(td_abs_errors, valid, dim=0)
        priorities = self.pri_eta * max_d + (1 - self.pri_eta) * mean_d       # [B]
        return loss, valid_td_abs_errors, priorities

    def value_scale(self, x):
        """r2d1.py:332-335."""
        return torch.sign(x) * (torch.sqrt(abs(x) + 1) - 1) + self.value_scale_eps * x

    def inv_value_scale(self, z):
        """r2d1.py:337-341."""
        return torch.sign(z) * (((torch.sqrt(1 + 4 * self.value_scale_eps * (abs(z) + 1 + self.value_scale_eps)) - 1) /
                                 (2 * self.value_scale_eps)) ** 2 - 1)
