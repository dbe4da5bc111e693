"""Offline evaluation in the master process (mirror of ``rlpyt/samplers/serial/collectors.py:12-66``
``SerialEvalCollector``): dedicated evaluation envs are stepped until ``max_T`` steps per env or
``max_trajectories`` completed trajectories; no samples are recorded, only TrajInfos returned.

Both samplers of this package evaluate this way.  The reference's parallel samplers fan evaluation out
to their worker processes (``rlpyt/samplers/parallel/gpu/collectors.py:129-176``); here evaluation is
kept off the training step loop on purpose - it is outside the accelerated path and
this keeps the worker protocol of the hot path untouched - at the price of stepping the evaluation envs
serially.  ``agent.step`` receives host tensors and answers with host tensors (the agents return
results on the device their inputs came from), so any agent of the package works on any device."""
import numpy as np
import torch

from rlpyt_b200.utils.buffer import buffer_from_example


class SerialEvalCollector:

    def __init__(self, envs, agent, TrajInfoCls, max_T, max_trajectories=None):
        self.envs, self.agent, self.TrajInfoCls = envs, agent, TrajInfoCls
        self.max_T, self.max_trajectories = max_T, max_trajectories

    def collect_evaluation(self, itr):
        envs = self.envs
        traj_infos = [self.TrajInfoCls() for _ in envs]
        completed = []
        observations = [env.reset() for env in envs]
        observation = buffer_from_example(observations[0], len(envs))
        for b, o in enumerate(observations):
            observation[b] = o
        action = buffer_from_example(envs[0].action_space.null_value(), len(envs))
        reward = np.zeros(len(envs), dtype="float32")
        obs_pyt, act_pyt, rew_pyt = (torch.from_numpy(x) for x in (observation, action, reward))
        self.agent.reset()
        self.agent.eval_mode(itr)
        for _t in range(self.max_T):
            step = self.agent.step(obs_pyt, act_pyt, rew_pyt)
            action[...] = step.action.cpu().numpy()                   # act_pyt aliases action
            for b, env in enumerate(envs):
                o, r, d, env_info = env.step(action[b])
                traj_infos[b].step(observation[b], action[b], r, d, None, env_info)
                if getattr(env_info, "traj_done", d):
                    completed.append(traj_infos[b].terminate(o))
                    traj_infos[b] = self.TrajInfoCls()
                    o = env.reset()
                if d:
                    action[b] = 0                                      # prev_action for the next step
                    r = 0
                    self.agent.reset_one(idx=b)
                observation[b] = o
                reward[b] = r
            if self.max_trajectories is not None and len(completed) >= self.max_trajectories:
                break
        return completed


def build_eval_collector(sampler, agent, seed):
    """The evaluation set-up shared by the samplers (``rlpyt/samplers/serial/sampler.py:68-80``)."""
    from rlpyt_b200.utils.seed import set_envs_seeds
    if not sampler.eval_n_envs or sampler.eval_n_envs <= 0:
        return None
    eval_env_kwargs = sampler.eval_env_kwargs if sampler.eval_env_kwargs is not None else sampler.env_kwargs
    eval_envs = [sampler.EnvCls(**eval_env_kwargs) for _ in range(sampler.eval_n_envs)]
    set_envs_seeds(eval_envs, seed)
    Cls = sampler.eval_CollectorCls or SerialEvalCollector
    return Cls(envs=eval_envs, agent=agent, TrajInfoCls=sampler.TrajInfoCls,
               max_T=max(1, int(sampler.eval_max_steps) // sampler.eval_n_envs),
               max_trajectories=sampler.eval_max_trajectories)
