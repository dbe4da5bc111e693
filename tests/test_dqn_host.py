"""Host-side DQN pieces that need no GPU: constructor defaults and schedules of ``DQN``
(rlpyt/algos/dqn/dqn.py:28-93,267-283), the epsilon-greedy distribution and agent mixin
(rlpyt/distributions/epsilon_greedy.py, rlpyt/agents/dqn/epsilon_greedy.py), the Q-network's
parameter names (state_dict interchange with the reference), target-network updates, and that the
loss refuses to run without the CUDA library (no CPU fallback in the product path)."""
from collections import namedtuple

import numpy as np
import pytest
import torch

from rlpyt_b200.agents.dqn.atari.atari_dqn_agent import AtariDqnAgent
from rlpyt_b200.algos.dqn.dqn import DQN
from rlpyt_b200.distributions.epsilon_greedy import EpsilonGreedy
from rlpyt_b200.models.dqn.atari_dqn_model import AtariDqnModel

Spaces = namedtuple("Spaces", "observation action")
SPACES = Spaces(namedtuple("O", "shape")((4, 84, 84)), namedtuple("A", "n")(6))

# parameter names of rlpyt/models/dqn/atari_dqn_model.py (Conv2dModel.conv.<i>, MlpModel.model.<i>)
REF_KEYS = ["conv.conv.0.weight", "conv.conv.0.bias", "conv.conv.2.weight", "conv.conv.2.bias", "conv.conv.4.weight",
            "conv.conv.4.bias", "head.model.0.weight", "head.model.0.bias", "head.model.2.weight", "head.model.2.bias"]
REF_DUELING_KEYS = ["head.advantage_bias", "head.advantage_hidden.model.0.weight", "head.advantage_hidden.model.0.bias",
                    "head.advantage_out.weight", "head.value.model.0.weight", "head.value.model.0.bias",
                    "head.value.model.2.weight", "head.value.model.2.bias"]


def test_model_matches_reference_layout():
    m = AtariDqnModel((4, 84, 84), 6)
    assert list(m.state_dict().keys()) == REF_KEYS
    # the reference pads the 2nd and 3rd layers (paddings [0,1,1]): 84x84 -> 20x20 -> 10x10 -> 10x10
    assert m.state_dict()["head.model.0.weight"].shape == (512, 64 * 10 * 10)
    assert sum(p.numel() for p in m.parameters()) == 3_358_374
    d = AtariDqnModel((4, 104, 80), 4, dueling=True)
    assert [k for k in d.state_dict().keys() if k.startswith("head.")] == REF_DUELING_KEYS
    q = d(torch.zeros(2, 3, 4, 104, 80, dtype=torch.uint8), None, None)
    assert q.shape == (2, 3, 4)
    assert m(torch.zeros(4, 84, 84, dtype=torch.uint8), None, None).shape == (6,)


def test_dueling_head_arithmetic():
    from rlpyt_b200.models.dqn.dueling import DuelingHeadModel
    torch.manual_seed(0)
    h = DuelingHeadModel(10, 8, 5)
    x = torch.randn(3, 10, requires_grad=True)
    q = h(x)
    adv = h.advantage(x)
    np.testing.assert_allclose(q.detach().numpy(),
                               (h.value(x) + adv - adv.mean(-1, keepdim=True)).detach().numpy(), rtol=1e-6)
    q.sum().backward()
    x2 = x.detach().clone().requires_grad_(True)
    (h.value(x2) + h.advantage(x2) - h.advantage(x2).mean(-1, keepdim=True)).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), x2.grad.numpy() * 2 ** -0.5, rtol=1e-5)   # scale_grad


def test_epsilon_greedy_sampling():
    d = EpsilonGreedy(dim=4, epsilon=0.0)
    q = torch.tensor([[0.1, 0.9, 0.9, 0.2], [3.0, 1.0, 2.0, 0.0]])
    assert d.sample(q).tolist() == [1, 0]                      # greedy; ties -> first index
    d.set_epsilon(1.0)
    torch.manual_seed(0)
    draws = torch.stack([d.sample(torch.zeros(1000, 4)) for _ in range(4)])
    assert draws.min() == 0 and draws.max() == 3
    assert abs(float((draws == 2).float().mean()) - 0.25) < 0.03
    d.set_epsilon(torch.tensor([0.0, 1.0]))                    # vector epsilon over the batch dim
    qq = torch.zeros(50, 2, 4)
    qq[..., 3] = 1.0
    s = d.sample(qq)
    assert torch.all(s[:, 0] == 3) and not torch.all(s[:, 1] == 3)
    assert d.to_onehot(torch.tensor([2])).tolist() == [[0.0, 0.0, 1.0, 0.0]]


def test_agent_epsilon_schedule_and_target_update():
    torch.manual_seed(1)
    agent = AtariDqnAgent(eps_init=1.0, eps_final=0.1, eps_eval=0.001)
    agent.initialize(SPACES)
    assert list(agent.state_dict().keys()) == ["model", "target"]
    for a, b in zip(agent.model.state_dict().values(), agent.target_model.state_dict().values()):
        assert torch.equal(a, b)
    agent.set_epsilon_itr_min_max(10, 110)
    for itr, want in [(0, 1.0), (10, 1.0), (60, 0.55), (110, 0.1), (500, 0.1)]:
        agent.sample_mode(itr)
        assert abs(agent.distribution.epsilon - want) < 1e-12
    agent.eval_mode(0)
    assert agent.distribution.epsilon == 1.0
    agent.eval_mode(5)
    assert agent.distribution.epsilon == 0.001
    with torch.no_grad():
        for p in agent.model.parameters():
            p.add_(1.0)
    agent.update_target(tau=0.25)                               # soft update: 0.25*new + 0.75*old
    for new, tgt in zip(agent.model.state_dict().values(), agent.target_model.state_dict().values()):
        np.testing.assert_allclose(tgt.numpy(), (new - 0.75).numpy(), rtol=1e-6, atol=1e-6)
    agent.update_target()
    for a, b in zip(agent.model.state_dict().values(), agent.target_model.state_dict().values()):
        assert torch.equal(a, b)
    step = agent.step(torch.zeros(3, 4, 84, 84, dtype=torch.uint8), torch.zeros(3, dtype=torch.long), torch.zeros(3))
    assert step.action.shape == (3,) and step.agent_info.q.shape == (3, 6)


def test_vector_epsilon_log_spaced():
    agent = AtariDqnAgent(eps_final=0.1, eps_final_min=0.001)
    agent.initialize(SPACES, global_B=8, env_ranks=[2, 3])
    want = torch.logspace(-3, -1, 8)[[2, 3]]
    np.testing.assert_allclose(agent.eps_final.numpy(), want.numpy(), rtol=1e-6)
    assert agent.eps_init.tolist() == [1.0, 1.0]


def test_dqn_constructor_defaults_and_schedules():
    algo = DQN(batch_size=32)
    assert algo.optim_kwargs == dict(eps=0.01 / 32) and algo.default_priority == 1.0 and algo.batch_size == 32
    assert algo.opt_info_fields == ("loss", "gradNorm", "tdAbsErr")
    algo = DQN(prioritized_replay=True, pri_beta_init=0.4, pri_beta_final=1.0)
    algo.min_itr_learn, algo.pri_beta_itr = 10, 110
    betas = []
    algo.replay_buffer = namedtuple("R", "set_beta")(betas.append)
    for itr in (0, 10, 60, 110, 111):
        algo.update_itr_hyperparams(itr)
    np.testing.assert_allclose(betas, [0.4, 0.4, 0.7, 1.0], rtol=1e-12)     # itr 111 > pri_beta_itr: unchanged


def test_loss_fails_loudly_without_cuda():
    """No CPU fallback: CPU tensors are rejected by the library wrapper."""
    from rlpyt_b200 import _lib
    from rlpyt_b200.algos.dqn import loss_ops
    q = torch.zeros(4, 3, requires_grad=True)
    with pytest.raises(_lib.B200LibraryError):
        loss_ops.dqn_loss(q, torch.zeros(4, 3), None, torch.zeros(4, dtype=torch.long), torch.zeros(4),
                          torch.zeros(4, dtype=torch.bool), None, 0.99, 1.0)


def test_grad_destination_claims_a_slot_once_per_zero_grad():
    """algos/optim.py ``grad_destination`` (host logic, CPU tensors): a parameter's slot of the flat gradient buffer is
    handed out once per ``zero_grad`` and only while ``.grad`` is empty; everything else gets a fresh tensor that autograd
    accumulates as usual; and autograd adopts the slot view without copying."""
    import torch
    from rlpyt_b200.algos.optim import grad_destination

    class Owner:
        _stamp, direct_grads = 1, True
    flat = torch.zeros(12)
    p = torch.nn.Parameter(torch.randn(2, 3))
    p._flat_grad = dict(owner=Owner, slot=flat[4:10], stamp=-1)
    p.grad = None
    a = grad_destination(p)
    assert a.data_ptr() == flat[4:10].data_ptr() and a.shape == p.shape
    b = grad_destination(p)                                   # second use in the same backward: accumulate the usual way
    assert b.data_ptr() != a.data_ptr() and b.shape == p.shape
    Owner._stamp = 2                                          # zero_grad
    p.grad = torch.zeros(2, 3)                                # ... but a gradient has already arrived from elsewhere
    assert grad_destination(p).data_ptr() != flat[4:10].data_ptr()
    p.grad = None
    Owner.direct_grads = False
    assert grad_destination(p).data_ptr() != flat[4:10].data_ptr()
    Owner.direct_grads = True
    assert grad_destination(p).data_ptr() == flat[4:10].data_ptr()
    q = torch.nn.Parameter(torch.randn(3))                    # a parameter no FlatAdam manages
    assert grad_destination(q).shape == q.shape

    class Mul(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            ctx.w = w
            ctx.save_for_backward(x)
            return x * w

        @staticmethod
        def backward(ctx, g):
            (x,) = ctx.saved_tensors
            dest = grad_destination(ctx.w)
            torch.mul(g, x, out=dest)
            return g * ctx.w, dest
    Owner._stamp = 3
    p.grad = None
    x = torch.randn(2, 3)
    Mul.apply(x, p).sum().backward()
    assert p.grad.data_ptr() == flat[4:10].data_ptr() and torch.equal(flat[4:10].view(2, 3), x)   # adopted, not copied
