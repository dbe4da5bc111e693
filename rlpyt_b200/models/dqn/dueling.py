"""Dueling head (mirror of ``rlpyt/models/dqn/dueling.py:7-46``; same parameter names)."""
import torch

from rlpyt_b200.models.mlp import MlpModel
from rlpyt_b200.models.utils import scale_grad


class DuelingHeadModel(torch.nn.Module):
    """Q = V + (A - mean(A)) with a shared advantage bias and gradient scaling into the trunk."""

    def __init__(self, input_size, hidden_sizes, output_size, grad_scale=2 ** (-1 / 2)):
        super().__init__()
        if isinstance(hidden_sizes, int):
            hidden_sizes = [hidden_sizes]
        self.advantage_hidden = MlpModel(input_size, hidden_sizes)
        self.advantage_out = torch.nn.Linear(hidden_sizes[-1], output_size, bias=False)
        self.advantage_bias = torch.nn.Parameter(torch.zeros(1))
        self.value = MlpModel(input_size, hidden_sizes, output_size=1)
        self._grad_scale = grad_scale

    def forward(self, input):
        x = scale_grad(input, self._grad_scale)
        advantage = self.advantage(x)
        value = self.value(x)
        return value + (advantage - advantage.mean(dim=-1, keepdim=True))

    def advantage(self, input):
        return self.advantage_out(self.advantage_hidden(input)) + self.advantage_bias
