"""Sampler interface handed to the runner (mirror of ``rlpyt/samplers/base.py:7-67``)."""
from rlpyt_b200.samplers.collections import BatchSpec, TrajInfo


class BaseSampler:

    alternating = False

    def __init__(self, EnvCls, env_kwargs, batch_T, batch_B, CollectorCls=None, max_decorrelation_steps=100,
                 TrajInfoCls=TrajInfo, eval_n_envs=0, eval_CollectorCls=None, eval_env_kwargs=None,
                 eval_max_steps=None, eval_max_trajectories=None):
        self.EnvCls = EnvCls
        self.env_kwargs = env_kwargs
        self.batch_T, self.batch_B = batch_T, batch_B
        self.CollectorCls = CollectorCls
        self.max_decorrelation_steps = max_decorrelation_steps
        self.TrajInfoCls = TrajInfoCls
        self.eval_n_envs = eval_n_envs
        self.eval_CollectorCls = eval_CollectorCls
        self.eval_env_kwargs = eval_env_kwargs
        self.eval_max_steps = None if eval_max_steps is None else int(eval_max_steps)
        self.eval_max_trajectories = None if eval_max_trajectories is None else int(eval_max_trajectories)
        self.batch_spec = BatchSpec(batch_T, batch_B)
        self.mid_batch_reset = getattr(CollectorCls, "mid_batch_reset", True)

    def initialize(self, *args, **kwargs):
        raise NotImplementedError

    def obtain_samples(self, itr):
        raise NotImplementedError

    def evaluate_agent(self, itr):
        raise NotImplementedError

    def shutdown(self):
        pass

    @property
    def batch_size(self):
        return self.batch_spec.size
