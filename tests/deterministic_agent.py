"""A policy whose outputs are exact integer/dyadic functions of its inputs, for collector parity tests: the same
arithmetic runs inside the reference's samplers (torch-CPU, tests/golden/make_golden.py) and inside this repo's
samplers (CUDA), so the recorded [T,B] batches must agree field by field - no sampling, no network rounding.

    action[b] = (sum(obs[b, 0, 0, :8]) + 3 * prev_action[b] + (prev_reward[b] != 0)) % A
    prob[b]   = one_hot(action[b]) * 0.75 + 0.25 / A         (exact in fp32 for A in {4, 5})
    value[b]  = obs[b, 1, 2, 3] / 4 + prev_reward[b] / 2     (exact)
"""
import torch


def policy(observation, prev_action, prev_reward, n_actions):
    o = observation.reshape((-1,) + tuple(observation.shape[-3:]))
    pa = prev_action.reshape(-1).to(torch.int64)
    pr = prev_reward.reshape(-1).to(torch.float32)
    s = o[:, 0, 0, :8].to(torch.int64).sum(-1)
    action = (s + 3 * pa + (pr != 0).to(torch.int64)) % n_actions
    prob = torch.nn.functional.one_hot(action, n_actions).to(torch.float32) * 0.75 + 0.25 / n_actions
    value = o[:, 1, 2, 3].to(torch.float32) / 4 + pr / 2
    lead = tuple(observation.shape[:-3])
    return action.reshape(lead), prob.reshape(lead + (n_actions,)), value.reshape(lead)


def make_agent_class(AgentStep, AgentInfo, DistInfo):
    """Duck-typed agent over the namedarraytuple classes of either package (reference or rlpyt_b200)."""

    class DeterministicAgent:
        recurrent = False
        alternating = False

        def __init__(self):
            self.device = torch.device("cpu")
            self.n_actions = None
            self._mode = None

        # ---- the surface the samplers touch (rlpyt/agents/base.py:59-216)
        def initialize(self, env_spaces, share_memory=False, global_B=1, env_ranks=None):
            self.n_actions = int(env_spaces.action.n)
            self.env_spaces = env_spaces

        def to_device(self, cuda_idx=None):
            if cuda_idx is not None:
                self.device = torch.device("cuda", cuda_idx)

        def data_parallel(self):
            pass

        def collector_initialize(self, global_B=1, env_ranks=None):
            pass

        def reset(self):
            pass

        def reset_one(self, idx):
            pass

        def sync_shared_memory(self):
            pass

        def toggle_alt(self):
            pass

        def sample_mode(self, itr):
            self._mode = "sample"

        def train_mode(self, itr):
            self._mode = "train"

        def eval_mode(self, itr):
            self._mode = "eval"

        def parameters(self):
            return []

        def state_dict(self):
            return {}

        @torch.no_grad()
        def step(self, observation, prev_action, prev_reward):
            action, prob, value = policy(observation, prev_action, prev_reward, self.n_actions)
            return AgentStep(action=action, agent_info=AgentInfo(dist_info=DistInfo(prob=prob), value=value))

        @torch.no_grad()
        def value(self, observation, prev_action, prev_reward):
            return policy(observation, prev_action, prev_reward, self.n_actions)[2]

    return DeterministicAgent
