"""Oracle: the serial CPU rollout loop (torch-CPU agent, per-env Python stepping).

Test infrastructure / CPU baseline only (see oracle/__init__.py).  Restates
``CpuResetCollector.collect_batch`` (rlpyt/samplers/parallel/cpu/collectors.py:25-65) as driven by
``SerialSampler.obtain_samples`` (rlpyt/samplers/serial/sampler.py:94-105) with
``CategoricalPgAgent.step`` (rlpyt/agents/pg/categorical.py:33-43: forward, ``torch.multinomial``
sample) on the oracle network of oracle/atari_ff.py.  Buffers follow rlpyt/samplers/buffer.py:28-45
(``[T+1,B]`` action / reward blocks whose ``[1:]`` / ``[:-1]`` views alias).
"""
import numpy as np
import torch

from oracle import atari_ff


class SerialRollout:

    def __init__(self, envs, state_dict, T, n_actions, agent_fn=None):
        """``agent_fn(obs, prev_action, prev_reward) -> (action, prob, value)`` (torch-CPU tensors) replaces the
        network + multinomial draw - used to pin this loop to the reference's SerialSampler under a deterministic
        policy (tests/golden/collector.npz, tests/test_oracle_collector.py)."""
        self.envs, self.sd, self.T, self.A = envs, state_dict, T, n_actions
        self.agent_fn = agent_fn
        B = len(envs)
        obs0 = np.stack([e.reset() for e in envs])
        self.observation = obs0.copy()
        self.prev_action = np.zeros(B, dtype=np.int64)
        self.prev_reward = np.zeros(B, dtype=np.float32)
        self.buf = dict(
            observation=np.zeros((T, B) + obs0.shape[1:], dtype=obs0.dtype),
            all_action=np.zeros((T + 1, B), dtype=np.int64),
            all_reward=np.zeros((T + 1, B), dtype=np.float32),
            done=np.zeros((T, B), dtype=bool),
            prob=np.zeros((T, B, n_actions), dtype=np.float32),
            value=np.zeros((T, B), dtype=np.float32),
            bootstrap_value=np.zeros((1, B), dtype=np.float32),
        )

    @torch.no_grad()
    def collect_batch(self, state_dict=None):
        sd = self.sd if state_dict is None else state_dict
        buf, T = self.buf, self.T
        buf["all_action"][0] = self.prev_action                                   # cpu/collectors.py:33-34
        buf["all_reward"][0] = self.prev_reward
        for t in range(T):
            buf["observation"][t] = self.observation                              # :37
            if self.agent_fn is not None:
                a_t, pi, v = self.agent_fn(torch.from_numpy(self.observation), torch.from_numpy(self.prev_action),
                                           torch.from_numpy(self.prev_reward))
                action = a_t.numpy()
            else:
                pi, v = atari_ff.forward(sd, torch.from_numpy(self.observation))       # categorical.py:37
                action = torch.multinomial(pi, num_samples=1).squeeze(-1).numpy()      # categorical.py:29
            for b, env in enumerate(self.envs):                                    # :41-54
                o, r, d, info = env.step(action[b])
                if getattr(info, "traj_done", d):
                    o = env.reset()
                self.observation[b] = o
                self.prev_reward[b] = r
                buf["done"][t, b] = d
            buf["all_action"][t + 1] = action                                      # :55
            buf["all_reward"][t + 1] = self.prev_reward                            # :56
            buf["prob"][t] = pi.numpy()
            buf["value"][t] = v.numpy()
            self.prev_action = action
        if self.agent_fn is not None:
            _, _, v = self.agent_fn(torch.from_numpy(self.observation), torch.from_numpy(self.prev_action),
                                    torch.from_numpy(self.prev_reward))
        else:
            _, v = atari_ff.forward(sd, torch.from_numpy(self.observation))            # :61-63
        buf["bootstrap_value"][0] = v.numpy()
        return buf
