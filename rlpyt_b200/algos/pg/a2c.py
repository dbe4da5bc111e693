"""Synchronous Advantage Actor-Critic (mirror of ``rlpyt/algos/pg/a2c.py:11-103``): one
full-batch gradient step per iteration, returns + loss + update all on the device."""
import torch

from rlpyt_b200.agents.base import AgentInputs
from rlpyt_b200.algos.optim import FlatAdam
from rlpyt_b200.algos.pg import loss_ops
from rlpyt_b200.algos.pg.base import PolicyGradientAlgo, OptInfo


class A2C(PolicyGradientAlgo):

    def __init__(self, discount=0.99, learning_rate=0.001, value_loss_coeff=0.5, entropy_loss_coeff=0.01,
                 OptimCls=FlatAdam, optim_kwargs=None, clip_grad_norm=1., initial_optim_state_dict=None,
                 gae_lambda=1, normalize_advantage=False):
        self.discount = discount
        self.learning_rate = learning_rate
        self.value_loss_coeff = value_loss_coeff
        self.entropy_loss_coeff = entropy_loss_coeff
        self.OptimCls = OptimCls
        self.optim_kwargs = dict() if optim_kwargs is None else optim_kwargs
        self.clip_grad_norm = clip_grad_norm
        self.initial_optim_state_dict = initial_optim_state_dict
        self.gae_lambda = gae_lambda
        self.normalize_advantage = normalize_advantage

    def initialize(self, *args, **kwargs):
        super().initialize(*args, **kwargs)
        self._batch_size = self.batch_spec.size

    def optimize_agent(self, itr, samples):
        """a2c.py:41-61."""
        if hasattr(self.agent, "update_obs_rms"):
            self.agent.update_obs_rms(samples.env.observation)
        self.optimizer.zero_grad()
        loss, entropy, perplexity = self.loss(samples)
        loss.backward()
        if isinstance(self.optimizer, FlatAdam):
            grad_norm = self.optimizer.clip_and_step(self.clip_grad_norm)
        else:
            grad_norm = torch.nn.utils.clip_grad_norm_(self.agent.parameters(), self.clip_grad_norm)
            self.optimizer.step()
        host = torch.stack([loss.detach().reshape(()), grad_norm.detach().reshape(()).to(loss.device),
                            entropy.reshape(()), perplexity.reshape(())]).cpu().tolist()
        self.update_counter += 1
        return OptInfo(loss=host[0], gradNorm=host[1], entropy=host[2], perplexity=host[3])

    def loss(self, samples):
        """a2c.py:63-103; everything on the device."""
        agent_inputs = AgentInputs(
            observation=self._on_device(samples.env.observation),
            prev_action=self._on_device(samples.agent.prev_action),
            prev_reward=self._on_device(samples.env.prev_reward),
        )
        if self.agent.recurrent:                                          # a2c.py:73-79
            from rlpyt_b200.utils.buffer import buffer_method, buffer_to
            init = buffer_to(samples.agent.agent_info.prev_rnn_state[0], device=self._device())       # T = 0: [B,N,H]
            init = buffer_method(buffer_method(init, "transpose", 0, 1), "contiguous")
            dist_info, value, _rnn_state = self.agent(*agent_inputs, init)
        else:
            dist_info, value = self.agent(*agent_inputs)
        return_, advantage, valid = self.process_returns(samples)
        action = self._on_device(samples.agent.action)
        loss, sc = loss_ops.a2c_loss(dist_info.prob, value, action, return_, advantage, valid,
                                     self.value_loss_coeff, self.entropy_loss_coeff)
        return loss, sc[1], sc[2]
