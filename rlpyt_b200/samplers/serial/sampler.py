"""In-process sampler (mirror of ``rlpyt/samplers/serial/sampler.py:10-109`` driving the loop of
``rlpyt/samplers/parallel/cpu/collectors.py:25-65`` ``CpuResetCollector``): environments are stepped
by the master itself; ``agent.step`` runs on the device and the batch is recorded in HBM by the
same step engine as the parallel sampler.  Used by BASELINE.json config 1 (A2C, T=5, B=8)."""
import numpy as np
import torch

from rlpyt_b200.samplers.base import BaseSampler
from rlpyt_b200.samplers.buffer import build_samples_buffer
from rlpyt_b200.samplers.collectors import DecorrelatingStartCollector
from rlpyt_b200.samplers.eval_collector import build_eval_collector
from rlpyt_b200.samplers.rollout import DeviceRollout
from rlpyt_b200.utils.seed import set_envs_seeds


class SerialSampler(BaseSampler):

    def initialize(self, agent, affinity=None, seed=None, bootstrap_value=False, traj_info_kwargs=None,
                   rank=0, world_size=1):
        B = self.batch_spec.B
        self.envs = [self.EnvCls(**self.env_kwargs) for _ in range(B)]
        set_envs_seeds(self.envs, seed)
        global_B = B * world_size
        env_ranks = list(range(rank * B, (rank + 1) * B))
        cuda_idx = (affinity or {}).get("cuda_idx", None)
        if cuda_idx is None:
            cuda_idx = torch.cuda.current_device()
        self.device = torch.device("cuda", cuda_idx)
        if not getattr(self, "_agent_preinitialized", False):   # the asynchronous samplers initialize the agent in async_initialize
            agent.initialize(self.envs[0].spaces, share_memory=False, global_B=global_B, env_ranks=env_ranks)
        self.agent = agent
        self.samples, self.host, examples = build_samples_buffer(
            agent, self.envs[0], self.batch_spec, bootstrap_value, device=self.device, share_host=False)
        if traj_info_kwargs:
            for k, v in traj_info_kwargs.items():
                setattr(self.TrajInfoCls, "_" + k, v)
        starter = DecorrelatingStartCollector(rank=0, envs=self.envs, env_info_np=self.host["env_info_np"],
                                              batch_T=self.batch_spec.T, TrajInfoCls=self.TrajInfoCls,
                                              step_buffer_np=self.host["step_np"], global_B=global_B,
                                              env_ranks=env_ranks)
        self.traj_infos = starter.start_envs(self.max_decorrelation_steps)
        agent.collector_initialize(global_B=global_B, env_ranks=env_ranks)
        agent.reset()
        agent.sample_mode(itr=0)
        self.rollout = DeviceRollout(self.samples, self.host, agent, self.device)
        self.rollout.in_action.copy_(self.host["step_pyt"].action)
        self.samples_pyt = self.samples
        self.eval_collector = build_eval_collector(self, agent, seed)
        return examples

    def obtain_samples(self, itr):
        """cpu/collectors.py:25-65: no input zeroing on done, only ``agent.reset_one``."""
        step, ro, T = self.host["step_np"], self.rollout, self.batch_spec.T
        env_info_np = self.host["env_info_np"]
        completed = []
        self.agent.sample_mode(itr)
        for t in range(T):
            ro.step(t, zero_inputs_on_done=False)
            for b, env in enumerate(self.envs):
                o, r, d, env_info = env.step(step.action[b])
                self.traj_infos[b].step(step.observation[b], step.action[b], r, d, None, env_info)
                if getattr(env_info, "traj_done", d):
                    completed.append(self.traj_infos[b].terminate(o))
                    self.traj_infos[b] = self.TrajInfoCls()
                    o = env.reset()
                if d:
                    self.agent.reset_one(idx=b)
                step.observation[b] = o
                step.reward[b] = r
                step.done[b] = d
                if env_info:
                    env_info_np[t, b] = env_info
        ro.finish()
        torch.cuda.current_stream(self.device).synchronize()
        ro.end_batch()
        return self.samples, completed

    def evaluate_agent(self, itr):
        """serial/sampler.py:107-109."""
        if self.eval_collector is None:
            raise RuntimeError("evaluate_agent needs eval_n_envs > 0 (and eval_max_steps) at construction")
        return self.eval_collector.collect_evaluation(itr)
