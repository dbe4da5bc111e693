"""rlpyt/agents/dqn/atari/atari_dqn_agent.py:7-10."""
from rlpyt_b200.agents.dqn.atari.mixin import AtariMixin
from rlpyt_b200.agents.dqn.dqn_agent import DqnAgent
from rlpyt_b200.models.dqn.atari_dqn_model import AtariDqnModel


class AtariDqnAgent(AtariMixin, DqnAgent):

    def __init__(self, ModelCls=AtariDqnModel, **kwargs):
        super().__init__(ModelCls=ModelCls, **kwargs)
