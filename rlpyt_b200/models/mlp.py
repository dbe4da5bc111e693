"""MLP block (mirror of ``rlpyt/models/mlp.py:5-47``; same submodule names so reference
state_dicts load: ``model.<i>.weight``)."""
import torch


class MlpModel(torch.nn.Module):
    """Linear(+nonlinearity) stack; a final plain Linear iff ``output_size`` is given."""

    def __init__(self, input_size, hidden_sizes, output_size=None, nonlinearity=torch.nn.ReLU):
        super().__init__()
        if hidden_sizes is None:
            hidden_sizes = []
        elif isinstance(hidden_sizes, int):
            hidden_sizes = [hidden_sizes]
        layers, width = [], input_size
        for h in hidden_sizes:
            layers += [torch.nn.Linear(width, h), nonlinearity()]
            width = h
        if output_size is not None:
            layers.append(torch.nn.Linear(width, output_size))
            width = output_size
        self.model = torch.nn.Sequential(*layers)
        self._output_size = width

    def forward(self, input):
        return self.model(input)

    @property
    def output_size(self):
        return self._output_size
