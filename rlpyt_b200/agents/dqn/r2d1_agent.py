"""Recurrent DQN agent (mirror of ``rlpyt/agents/dqn/r2d1_agent.py:11-59``).  As for the other agents the network
outputs stay on the device; ``step`` answers on the device its observation came from, and stores the recurrent state
it STARTED from, transposed to [B,N,H], in ``agent_info.prev_rnn_state`` (what the sequence replay keeps)."""
import torch

from rlpyt_b200.agents.base import AgentStep, AlternatingRecurrentAgentMixin, RecurrentAgentMixin
from rlpyt_b200.agents.dqn.dqn_agent import DqnAgent
from rlpyt_b200.utils.buffer import buffer_to
from rlpyt_b200.utils.collections import namedarraytuple

AgentInfo = namedarraytuple("AgentInfo", ["q", "prev_rnn_state"])


def _map(state, fn):
    return None if state is None else type(state)(*(fn(x) for x in state))


class R2d1AgentBase(DqnAgent):

    def __call__(self, observation, prev_action, prev_reward, init_rnn_state):
        """(q, next_rnn_state) with grad; ``init_rnn_state`` already [N,B,H] (r2d1_agent.py:17-24)."""
        inputs = self._model_inputs(observation, prev_action, prev_reward)
        return self.model(*inputs, buffer_to(init_rnn_state, device=self.device))

    @torch.no_grad()
    def step(self, observation, prev_action, prev_reward):
        """r2d1_agent.py:26-45."""
        home = observation.device
        q, rnn_state = self.model(*self._model_inputs(observation, prev_action, prev_reward), self.prev_rnn_state)
        action = self.distribution.sample(q)
        prev_rnn_state = self.prev_rnn_state or _map(rnn_state, torch.zeros_like)
        prev_rnn_state = _map(prev_rnn_state, lambda x: x.transpose(0, 1))          # [N,B,H] -> [B,N,H] for storage
        agent_info = AgentInfo(q=q, prev_rnn_state=prev_rnn_state)
        self.advance_rnn_state(rnn_state)
        if home != self.device:
            action, agent_info = buffer_to((action, agent_info), device=home)
        return AgentStep(action=action, agent_info=agent_info)

    @torch.no_grad()
    def target(self, observation, prev_action, prev_reward, init_rnn_state):
        """r2d1_agent.py:47-53."""
        inputs = self._model_inputs(observation, prev_action, prev_reward)
        return self.target_model(*inputs, buffer_to(init_rnn_state, device=self.device))


class R2d1Agent(RecurrentAgentMixin, R2d1AgentBase):
    pass


class R2d1AlternatingAgent(AlternatingRecurrentAgentMixin, R2d1AgentBase):
    pass
