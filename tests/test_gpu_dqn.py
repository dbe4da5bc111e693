"""GPU parity of the DQN path: the fused loss kernel against the reference's
``DQN.loss`` outputs in tests/golden/dqn.npz (TD-error priorities bit-exact, loss and gradient 1e-5
relative), the same through ``DQN.loss`` with a stub agent, and the learner end to end on the device
replay (one update compared with the CPU oracle on the very batch the device replay sampled)."""
from collections import namedtuple

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dqn_cases import CASES, load_case  # noqa: E402

RTOL = 1e-5


def _c(x):
    return None if x is None else torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _check(c, loss, td, grad):
    assert np.array_equal(td.cpu().numpy(), c["td_abs_errors"])             # priorities: bit-exact
    np.testing.assert_allclose(float(loss.detach()), c["loss"], rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(grad.cpu().numpy(), c["grad_qs"], rtol=RTOL, atol=1e-9)


@pytest.mark.parametrize("name", CASES)
def test_fused_dqn_loss_vs_reference(golden, name):
    from rlpyt_b200.algos.dqn import loss_ops
    c = load_case(golden("dqn"), name)
    qs = _c(c["qs"]).requires_grad_(True)
    loss, td = loss_ops.dqn_loss(qs, _c(c["target_qs"]), _c(c["next_qs"]) if c["double"] else None, _c(c["action"]),
                                 _c(c["return_"]), _c(c["done_n"]), _c(c["is_weights"]) if c["pri"] else None,
                                 c["discount"] ** c["n_step"], c["clip"])
    loss.backward()
    _check(c, loss, td, qs.grad)
    assert not td.requires_grad and td.shape == (c["qs"].shape[0],)


def test_upstream_gradient_scales(golden):
    from rlpyt_b200.algos.dqn import loss_ops
    c = load_case(golden("dqn"), "dqn_double_pri")
    qs = _c(c["qs"]).requires_grad_(True)
    loss, _ = loss_ops.dqn_loss(qs, _c(c["target_qs"]), _c(c["next_qs"]), _c(c["action"]), _c(c["return_"]),
                                _c(c["done_n"]), _c(c["is_weights"]), c["discount"] ** c["n_step"], c["clip"])
    (2.5 * loss).backward()
    np.testing.assert_allclose(qs.grad.cpu().numpy(), 2.5 * c["grad_qs"], rtol=RTOL, atol=1e-9)


def test_c_abi_argument_checks():
    from rlpyt_b200 import _lib
    q = torch.zeros(4, 3, device="cuda")
    with pytest.raises(_lib.B200LibraryError):
        _lib.call("rl_dqn_loss_f32", _lib.ptr(q), None, None, None, None, None, None, 4, 3, 0.99, 1.0, None, None,
                  None, None, _lib.stream())
    with pytest.raises(_lib.B200LibraryError):      # A above the kernel's limit
        big = torch.zeros(2, 65, device="cuda")
        a, r, d = torch.zeros(2, dtype=torch.int64, device="cuda"), torch.zeros(2, device="cuda"), \
            torch.zeros(2, dtype=torch.uint8, device="cuda")
        sc, td, ws = torch.zeros(2, device="cuda"), torch.zeros(2, device="cuda"), \
            torch.zeros(4, dtype=torch.float64, device="cuda")
        _lib.call("rl_dqn_loss_f32", _lib.ptr(big), _lib.ptr(big), None, _lib.ptr(a), _lib.ptr(r), _lib.ptr(d), None,
                  2, 65, 0.99, 1.0, _lib.ptr(sc), _lib.ptr(td), None, _lib.ptr(ws), _lib.stream())


class StubAgent:
    """Fixed network outputs on the device (observation[0] tags which input set is evaluated)."""

    def __init__(self, qs, next_qs, target_qs):
        self.qs, self.next_qs, self.target_qs = qs, next_qs, target_qs

    def __call__(self, observation, prev_action, prev_reward):
        return self.qs if int(observation[0]) == 0 else self.next_qs

    def target(self, observation, prev_action, prev_reward):
        return self.target_qs


@pytest.mark.parametrize("name", CASES)
def test_dqn_loss_method_vs_reference(golden, name):
    """``DQN.loss(samples)`` - the reference's signature - on device tensors."""
    from rlpyt_b200.agents.base import AgentInputs
    from rlpyt_b200.algos.dqn.dqn import DQN
    c = load_case(golden("dqn"), name)
    N = c["qs"].shape[0]
    algo = DQN(discount=c["discount"], delta_clip=c["clip"], n_step_return=c["n_step"], double_dqn=c["double"],
               prioritized_replay=c["pri"])
    algo.mid_batch_reset = True
    qs = _c(c["qs"]).requires_grad_(True)
    algo.agent = StubAgent(qs, _c(c["next_qs"]), _c(c["target_qs"]))
    S = namedtuple("S", "agent_inputs action return_ done done_n target_inputs is_weights")
    z = torch.zeros(N, device="cuda")
    samples = S(AgentInputs(z, z, z), _c(c["action"]), _c(c["return_"]), _c(c["done_n"]), _c(c["done_n"]),
                AgentInputs(z + 1, z, z), _c(c["is_weights"]))
    loss, td = algo.loss(samples)
    loss.backward()
    _check(c, loss, td, qs.grad)


# ----------------------------------------------------------------------------------------------- learner
Spaces = namedtuple("Spaces", "observation action")
IMG, A = (4, 84, 84), 6


def _samples(seed, T, B):
    from rlpyt_b200.samplers.collections import AgentSamples, EnvSamples, Samples
    from rlpyt_b200.agents.dqn.dqn_agent import AgentInfo as QInfo
    g = torch.Generator(device="cuda").manual_seed(seed)
    obs = torch.randint(0, 256, (T, B) + IMG, dtype=torch.uint8, device="cuda", generator=g)
    act = torch.randint(0, A, (T + 1, B), device="cuda", generator=g)
    rew = torch.randn(T + 1, B, device="cuda", generator=g)
    done = torch.rand(T, B, device="cuda", generator=g) < 0.05
    q = torch.zeros(T, B, A, device="cuda")
    return Samples(agent=AgentSamples(act[1:], act[:-1], QInfo(q=q)),
                   env=EnvSamples(obs, rew[1:], rew[:-1], done, None))


def _make(prioritized, double, n_step, batch_size=32):
    from rlpyt_b200.agents.dqn.atari.atari_dqn_agent import AtariDqnAgent
    from rlpyt_b200.algos.dqn.dqn import DQN
    from rlpyt_b200.samplers.collections import BatchSpec
    torch.manual_seed(0)
    agent = AtariDqnAgent()
    agent.initialize(Spaces(namedtuple("O", "shape")(IMG), namedtuple("Ac", "n")(A)))
    agent.to_device(0)
    T, B = 8, 4
    algo = DQN(batch_size=batch_size, min_steps_learn=2 * T * B, replay_size=64 * B, replay_ratio=32,
               target_update_interval=3, n_step_return=n_step, double_dqn=double, prioritized_replay=prioritized)
    examples = dict(observation=np.zeros(IMG, np.uint8), action=np.int64(0), reward=np.float32(0),
                    done=np.bool_(False))
    algo.initialize(agent, n_itr=100, batch_spec=BatchSpec(T, B), mid_batch_reset=True, examples=examples)
    return agent, algo, T, B


@pytest.mark.parametrize("prioritized,double,n_step", [(True, True, 3), (False, False, 1)])
def test_dqn_learner_end_to_end(prioritized, double, n_step):
    """Plumbing and invariants: warm-up iterations return empty OptInfo, later ones run
    ``updates_per_optimize`` updates, the target network follows at ``target_update_interval``, the
    sum-tree root equals the sum of its leaves after priority updates, parameters move."""
    agent, algo, T, B = _make(prioritized, double, n_step)
    assert algo.updates_per_optimize == 32 and algo.min_itr_learn == 2
    np.random.seed(3)
    agent.train_mode(0)
    w0 = [p.detach().clone() for p in agent.parameters()]
    for itr in range(4):
        info = algo.optimize_agent(itr, _samples(itr, T, B))
        if itr < algo.min_itr_learn:
            assert info.loss == [] and info.gradNorm == [] and info.tdAbsErr == []
        else:
            assert len(info.loss) == len(info.gradNorm) == 32 and len(info.tdAbsErr) == 32 * 4
            assert np.all(np.isfinite(info.loss)) and np.all(np.isfinite(info.gradNorm))
            assert np.all(np.asarray(info.tdAbsErr) >= 0) and np.all(np.asarray(info.tdAbsErr) <= 1.0)
    assert algo.update_counter == 64
    assert any(not torch.equal(a, b) for a, b in zip(w0, agent.parameters()))
    # 63 = last multiple of 3 <= 64: the target is one update behind the online network
    assert any(not torch.equal(a, b) for a, b in
               zip(agent.model.state_dict().values(), agent.target_model.state_dict().values()))
    if prioritized:
        tree = algo.replay_buffer.priority_tree
        nodes = tree.tree.cpu().numpy()
        leaves = nodes[tree.low_idx:tree.high_idx]
        np.testing.assert_allclose(nodes[0], leaves.sum(), rtol=1e-12)
        assert leaves.max() <= 1.0 + 1e-12                       # delta_clip ** alpha bound on priorities


def test_dqn_update_matches_cpu_oracle():
    """One update on the batch the device replay sampled, recomputed on the CPU with the same weights
    (torch-CPU network + oracle/dqn_loss.py): loss, priorities and the head gradients agree to the
    fp32 conv/GEMM rounding of the two networks."""
    from oracle.dqn_loss import dqn_loss as oracle_loss
    from rlpyt_b200.models.dqn.atari_dqn_model import AtariDqnModel
    agent, algo, T, B = _make(True, True, 3)
    np.random.seed(4)
    for itr in range(3):
        algo.replay_buffer.append_samples(algo.samples_to_buffer(_samples(10 + itr, T, B)))
    agent.train_mode(0)
    batch = algo.replay_buffer.sample_batch(32)
    algo.optimizer.zero_grad()
    loss, td = algo.loss(batch)
    loss.backward()
    # CPU side: same weights (online == target at start)
    cpu = AtariDqnModel(IMG, A)
    cpu.load_state_dict({k: v.detach().cpu() for k, v in agent.model.state_dict().items()})
    to = lambda x: x.detach().cpu()
    qs = cpu(to(batch.agent_inputs.observation), None, None)
    with torch.no_grad():
        tq = cpu(to(batch.target_inputs.observation), None, None)
    o_loss, o_td, o_grad = oracle_loss(qs.detach(), tq, tq.clone(), to(batch.action), to(batch.return_),
                                       to(batch.done_n), to(batch.is_weights), 0.99, 3, 1.0)
    np.testing.assert_allclose(float(loss), float(o_loss), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(td.cpu().numpy(), o_td.numpy(), rtol=1e-4, atol=1e-5)
    qs.backward(o_grad)
    got = agent.model.head.model[2].weight.grad.cpu().numpy()
    want = cpu.head.model[2].weight.grad.numpy()
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-5 * float(np.abs(want).max()))
