"""Epsilon schedule shared by the epsilon-greedy agents (mirror of
``rlpyt/agents/dqn/epsilon_greedy.py:12-140``: linear ramp between ``eps_itr_min`` and ``eps_itr_max``,
evaluation epsilon, optional log-spaced per-environment epsilons)."""
import torch


class EpsilonGreedyAgentMixin:

    def __init__(self, eps_init=1, eps_final=0.01, eps_final_min=None, eps_itr_min=50, eps_itr_max=1000,
                 eps_eval=0.001, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.eps_init, self.eps_final, self.eps_final_min = eps_init, eps_final, eps_final_min
        self.eps_itr_min, self.eps_itr_max, self.eps_eval = eps_itr_min, eps_itr_max, eps_eval
        self._eps_final_scalar = eps_final
        self._eps_init_scalar = eps_init
        self.eps_sample = eps_init

    def collector_initialize(self, global_B=1, env_ranks=None):
        if env_ranks is not None:
            self.make_vec_eps(global_B, env_ranks)

    def make_vec_eps(self, global_B, env_ranks):
        """Log-spaced final epsilons, this rank's slice (epsilon_greedy.py:50-66)."""
        if self.eps_final_min is not None and self.eps_final_min != self._eps_final_scalar:
            if self.alternating:
                assert global_B % 2 == 0
                global_B = global_B // 2
                env_ranks = list(set([i // 2 for i in env_ranks]))
            self.eps_init = self._eps_init_scalar * torch.ones(len(env_ranks))
            global_eps_final = torch.logspace(torch.log10(torch.tensor(self.eps_final_min)),
                                              torch.log10(torch.tensor(self._eps_final_scalar)), global_B)
            self.eps_final = global_eps_final[env_ranks]
        self.eps_sample = self.eps_init

    def set_epsilon_itr_min_max(self, eps_itr_min, eps_itr_max):
        self.eps_itr_min, self.eps_itr_max = eps_itr_min, eps_itr_max

    def set_sample_epsilon_greedy(self, epsilon):
        self.distribution.set_epsilon(epsilon)

    def sample_mode(self, itr):
        """Anneal and install the sampling epsilon (epsilon_greedy.py:101-112)."""
        super().sample_mode(itr)
        itr_min, itr_max = self.eps_itr_min, self.eps_itr_max
        if itr <= itr_max:
            prog = min(1, max(0, itr - itr_min) / (itr_max - itr_min))
            self.eps_sample = prog * self.eps_final + (1 - prog) * self.eps_init
        self.distribution.set_epsilon(self.eps_sample)

    def eval_mode(self, itr):
        super().eval_mode(itr)
        self.distribution.set_epsilon(self.eps_eval if itr > 0 else 1.)
