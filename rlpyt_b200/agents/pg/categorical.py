"""Policy-gradient agent with a categorical action distribution (mirror of
``rlpyt/agents/pg/categorical.py:11-51``)."""
import torch

from rlpyt_b200.agents.base import AgentStep, AlternatingRecurrentAgentMixin, BaseAgent, RecurrentAgentMixin
from rlpyt_b200.agents.pg.base import AgentInfo, AgentInfoRnn
from rlpyt_b200.distributions.categorical import Categorical, DistInfo
from rlpyt_b200.utils.buffer import buffer_func, buffer_method, buffer_to
from rlpyt_b200.utils.gather import HostMappedFrames, LazyRows


class CategoricalPgAgent(BaseAgent):
    """The model maps (observation, one-hot prev_action, prev_reward) -> (pi, value)."""

    def initialize(self, env_spaces, share_memory=False, global_B=1, env_ranks=None):
        super().initialize(env_spaces, share_memory, global_B=global_B, env_ranks=env_ranks)
        self.distribution = Categorical(dim=env_spaces.action.n)

    def _model_inputs(self, observation, prev_action, prev_reward):
        prev_action = self.distribution.to_onehot(prev_action)  # categorical.py:21,35
        if isinstance(observation, (LazyRows, HostMappedFrames)):  # un-gathered minibatch rows / frames the first layer streams in itself
            return (observation,) + buffer_to((prev_action, prev_reward), device=self.device)
        return buffer_to((observation, prev_action, prev_reward), device=self.device)

    def __call__(self, observation, prev_action, prev_reward):
        """-> (DistInfo(prob), value); differentiable; outputs STAY on ``self.device``
        (reference: moved to cpu, categorical.py:25)."""
        pi, value = self.model(*self._model_inputs(observation, prev_action, prev_reward))
        return DistInfo(prob=pi), value

    @torch.no_grad()
    def step(self, observation, prev_action, prev_reward):
        """-> AgentStep(action, AgentInfo(dist_info, value)) on the device of ``observation``
        (categorical.py:33-43)."""
        home = observation.device
        inputs = self._model_inputs(observation, prev_action, prev_reward)
        if hasattr(self.model, "forward_step"):      # heads + softmax + draw fused into one kernel on the device
            pi, value, action = self.model.forward_step(*inputs, self.distribution)
            dist_info = DistInfo(prob=pi)
        else:
            pi, value = self.model(*inputs)
            dist_info = DistInfo(prob=pi)
            action = self.distribution.sample(dist_info)
        agent_info = AgentInfo(dist_info=dist_info, value=value)
        if home != self.device:
            action, agent_info = buffer_to((action, agent_info), device=home)
        return AgentStep(action=action, agent_info=agent_info)

    @torch.no_grad()
    def value(self, observation, prev_action, prev_reward):
        """Bootstrap value of the final observation (categorical.py:45-51)."""
        home = observation.device
        _pi, value = self.model(*self._model_inputs(observation, prev_action, prev_reward))
        return value if home == self.device else value.to(home)


class RecurrentCategoricalPgAgentBase(BaseAgent):
    """Recurrent policy (mirror of ``rlpyt/agents/pg/categorical.py:54-98``): the model maps (observation, one-hot
    prev_action, prev_reward, rnn_state) -> (pi, value, next_rnn_state); ``step`` records the state it STARTED from,
    transposed to ``[B,N,H]`` so that the sample buffer can be sliced over B."""

    def __call__(self, observation, prev_action, prev_reward, init_rnn_state):
        """``init_rnn_state`` already ``[N,B,H]``; outputs stay on the device."""
        prev_action = self.distribution.to_onehot(prev_action)
        inputs = buffer_to((observation, prev_action, prev_reward, init_rnn_state), device=self.device)
        pi, value, next_rnn_state = self.model(*inputs)
        return DistInfo(prob=pi), value, next_rnn_state

    def initialize(self, env_spaces, share_memory=False, global_B=1, env_ranks=None):
        super().initialize(env_spaces, share_memory, global_B=global_B, env_ranks=env_ranks)
        self.distribution = Categorical(dim=env_spaces.action.n)

    @torch.no_grad()
    def step(self, observation, prev_action, prev_reward):
        home = observation.device
        prev_action = self.distribution.to_onehot(prev_action)
        inputs = buffer_to((observation, prev_action, prev_reward), device=self.device)
        pi, value, rnn_state = self.model(*inputs, self.prev_rnn_state)
        dist_info = DistInfo(prob=pi)
        action = self.distribution.sample(dist_info)
        prev_rnn_state = self.prev_rnn_state or buffer_func(rnn_state, torch.zeros_like)   # buffers cannot hold None
        prev_rnn_state = buffer_method(prev_rnn_state, "transpose", 0, 1)                  # [N,B,H] -> [B,N,H]
        agent_info = AgentInfoRnn(dist_info=dist_info, value=value, prev_rnn_state=prev_rnn_state)
        if home != self.device:
            action, agent_info = buffer_to((action, agent_info), device=home)
        self.advance_rnn_state(rnn_state)
        return AgentStep(action=action, agent_info=agent_info)

    @torch.no_grad()
    def value(self, observation, prev_action, prev_reward):
        home = observation.device
        prev_action = self.distribution.to_onehot(prev_action)
        inputs = buffer_to((observation, prev_action, prev_reward), device=self.device)
        _pi, value, _rnn_state = self.model(*inputs, self.prev_rnn_state)
        return value if home == self.device else value.to(home)


class RecurrentCategoricalPgAgent(RecurrentAgentMixin, RecurrentCategoricalPgAgentBase):
    pass


class AlternatingRecurrentCategoricalPgAgent(AlternatingRecurrentAgentMixin, RecurrentCategoricalPgAgentBase):
    pass
