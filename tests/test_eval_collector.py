"""In-process evaluation (``rlpyt/samplers/serial/collectors.py:12-66`` semantics) on the CPU with a
small categorical agent on CartPole: stop conditions, trajectory bookkeeping, reset handling."""
from collections import namedtuple

import numpy as np
import torch

from rlpyt_b200.agents.pg.categorical import CategoricalPgAgent
from rlpyt_b200.envs.cartpole import CartPoleEnv
from rlpyt_b200.samplers.collections import TrajInfo
from rlpyt_b200.samplers.eval_collector import SerialEvalCollector, build_eval_collector


class TinyPgModel(torch.nn.Module):

    def __init__(self, obs_dim, n_actions):
        super().__init__()
        self.pi = torch.nn.Linear(obs_dim, n_actions)
        self.v = torch.nn.Linear(obs_dim, 1)

    def forward(self, observation, prev_action, prev_reward):
        x = observation.float()
        return torch.softmax(self.pi(x), dim=-1), self.v(x).squeeze(-1)


class TinyAgent(CategoricalPgAgent):

    def make_env_to_model_kwargs(self, env_spaces):
        return dict(obs_dim=env_spaces.observation.shape[0], n_actions=env_spaces.action.n)


def _agent():
    torch.manual_seed(0)
    agent = TinyAgent(ModelCls=TinyPgModel)
    agent.initialize(CartPoleEnv().spaces)
    return agent


def test_eval_stops_at_max_T_and_counts_trajectories():
    agent = _agent()
    envs = [CartPoleEnv() for _ in range(3)]
    for i, e in enumerate(envs):
        e.seed(i)
    col = SerialEvalCollector(envs, agent, TrajInfo, max_T=200)
    infos = col.collect_evaluation(itr=1)
    assert agent._mode == "eval"
    assert len(infos) >= 3                                   # a random-ish policy drops the pole within ~30 steps
    assert all(info.Length >= 1 and info.Return == info.Length for info in infos)   # CartPole: reward 1 per step
    assert sum(info.Length for info in infos) <= 3 * 200


def test_eval_stops_at_max_trajectories():
    agent = _agent()
    envs = [CartPoleEnv() for _ in range(4)]
    for i, e in enumerate(envs):
        e.seed(10 + i)
    col = SerialEvalCollector(envs, agent, TrajInfo, max_T=10_000, max_trajectories=5)
    infos = col.collect_evaluation(itr=0)
    assert 5 <= len(infos) <= 5 + len(envs) - 1             # the step that reaches the bound may finish several


def test_build_eval_collector_from_sampler_fields():
    S = namedtuple("S", "eval_n_envs eval_env_kwargs env_kwargs EnvCls eval_CollectorCls TrajInfoCls eval_max_steps "
                        "eval_max_trajectories")
    agent = _agent()
    assert build_eval_collector(S(0, None, {}, CartPoleEnv, None, TrajInfo, 100, None), agent, seed=0) is None
    col = build_eval_collector(S(2, None, {}, CartPoleEnv, None, TrajInfo, 101, 7), agent, seed=3)
    assert isinstance(col, SerialEvalCollector) and len(col.envs) == 2 and col.max_T == 50 and col.max_trajectories == 7
    infos = col.collect_evaluation(itr=2)
    assert isinstance(infos, list)
