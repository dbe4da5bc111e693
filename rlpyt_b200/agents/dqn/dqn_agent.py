"""DQN agent: online + target Q-networks and epsilon-greedy stepping (mirror of
``rlpyt/agents/dqn/dqn_agent.py:17-81``).  As for the policy-gradient agents, network outputs stay on
the device (the reference copies q to the CPU, :28,63,75) and ``step`` answers on the device its
observation came from."""
import torch

from rlpyt_b200.agents.base import AgentStep, BaseAgent
from rlpyt_b200.agents.dqn.epsilon_greedy import EpsilonGreedyAgentMixin
from rlpyt_b200.distributions.epsilon_greedy import EpsilonGreedy
from rlpyt_b200.models.utils import update_state_dict
from rlpyt_b200.utils.buffer import buffer_to
from rlpyt_b200.utils.collections import namedarraytuple

AgentInfo = namedarraytuple("AgentInfo", "q")


class DqnAgent(EpsilonGreedyAgentMixin, BaseAgent):

    def _model_inputs(self, observation, prev_action, prev_reward):
        prev_action = self.distribution.to_onehot(prev_action)
        return buffer_to((observation, prev_action, prev_reward), device=self.device)

    def __call__(self, observation, prev_action, prev_reward):
        """Q-values with grad, on ``self.device``."""
        return self.model(*self._model_inputs(observation, prev_action, prev_reward))

    def initialize(self, env_spaces, share_memory=False, global_B=1, env_ranks=None):
        """dqn_agent.py:30-46: online and target networks from the same state."""
        init = self.initial_model_state_dict
        self.initial_model_state_dict = None      # the base class must not try to load the {model,target} dict
        super().initialize(env_spaces, share_memory, global_B=global_B, env_ranks=env_ranks)
        self.target_model = self.ModelCls(**self.env_model_kwargs, **self.model_kwargs)
        if init is not None:
            self.model.load_state_dict(init["model"])
            self.target_model.load_state_dict(init["model"])
        else:
            self.target_model.load_state_dict(self.model.state_dict())
        self.distribution = EpsilonGreedy(dim=env_spaces.action.n)
        if env_ranks is not None:
            self.make_vec_eps(global_B, env_ranks)

    def to_device(self, cuda_idx=None):
        super().to_device(cuda_idx)
        self.target_model.to(self.device)

    def state_dict(self):
        return dict(model=self.model.state_dict(), target=self.target_model.state_dict())

    def load_state_dict(self, state_dict):
        self.model.load_state_dict(state_dict["model"])
        self.target_model.load_state_dict(state_dict["target"])

    @torch.no_grad()
    def step(self, observation, prev_action, prev_reward):
        """dqn_agent.py:55-67."""
        home = observation.device
        q = self.model(*self._model_inputs(observation, prev_action, prev_reward))
        action = self.distribution.sample(q)
        agent_info = AgentInfo(q=q)
        if home != self.device:
            action, agent_info = buffer_to((action, agent_info), device=home)
        return AgentStep(action=action, agent_info=agent_info)

    @torch.no_grad()
    def target(self, observation, prev_action, prev_reward):
        """Target-network Q-values (dqn_agent.py:69-75), on ``self.device``."""
        return self.target_model(*self._model_inputs(observation, prev_action, prev_reward))

    def update_target(self, tau=1):
        update_state_dict(self.target_model, self.model.state_dict(), tau)
