"""Policy-gradient agent with a categorical action distribution (mirror of
``rlpyt/agents/pg/categorical.py:11-51``)."""
import torch

from rlpyt_b200.agents.base import AgentStep, BaseAgent
from rlpyt_b200.agents.pg.base import AgentInfo
from rlpyt_b200.distributions.categorical import Categorical, DistInfo
from rlpyt_b200.utils.buffer import buffer_to
from rlpyt_b200.utils.gather import LazyRows


class CategoricalPgAgent(BaseAgent):
    """The model maps (observation, one-hot prev_action, prev_reward) -> (pi, value)."""

    def initialize(self, env_spaces, share_memory=False, global_B=1, env_ranks=None):
        super().initialize(env_spaces, share_memory, global_B=global_B, env_ranks=env_ranks)
        self.distribution = Categorical(dim=env_spaces.action.n)

    def _model_inputs(self, observation, prev_action, prev_reward):
        prev_action = self.distribution.to_onehot(prev_action)  # categorical.py:21,35
        if isinstance(observation, LazyRows):  # un-gathered minibatch rows, already on the device
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
        pi, value = self.model(*self._model_inputs(observation, prev_action, prev_reward))
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
