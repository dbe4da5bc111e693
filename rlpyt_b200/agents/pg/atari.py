"""Atari policy-gradient agents (mirror of ``rlpyt/agents/pg/atari.py:9-37``)."""
from rlpyt_b200.agents.pg.categorical import (AlternatingRecurrentCategoricalPgAgent, CategoricalPgAgent,
                                              RecurrentCategoricalPgAgent)
from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
from rlpyt_b200.models.pg.atari_lstm_model import AtariLstmModel


class AtariMixin:
    """Environment interface -> model kwargs (image shape, number of actions)."""

    def make_env_to_model_kwargs(self, env_spaces):
        return dict(image_shape=env_spaces.observation.shape, output_size=env_spaces.action.n)


class AtariFfAgent(AtariMixin, CategoricalPgAgent):

    def __init__(self, ModelCls=AtariFfModel, **kwargs):
        super().__init__(ModelCls=ModelCls, **kwargs)


class AtariLstmAgent(AtariMixin, RecurrentCategoricalPgAgent):

    def __init__(self, ModelCls=AtariLstmModel, **kwargs):
        super().__init__(ModelCls=ModelCls, **kwargs)


class AlternatingAtariLstmAgent(AtariMixin, AlternatingRecurrentCategoricalPgAgent):

    def __init__(self, ModelCls=AtariLstmModel, **kwargs):
        super().__init__(ModelCls=ModelCls, **kwargs)
