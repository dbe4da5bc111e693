"""Device-side Categorical.sample (csrc/categorical.cu) against the oracle: the inverse-CDF rule with injected
uniforms is bit-exact, the in-kernel Philox4x32-10 stream equals oracle/philox.py, the call counter lives on the
device (fresh draws under CUDA-graph replay), and the empirical frequencies follow the probabilities."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _probs(n, a, seed):
    rng = np.random.default_rng(seed)
    p = rng.dirichlet(np.ones(a) * 0.5, n).astype(np.float32)
    p[::7, 0] = 0.0                                       # zero-probability actions
    p[3::11] = np.eye(a, dtype=np.float32)[rng.integers(0, a, len(p[3::11]))]      # deterministic rows
    return p


@pytest.mark.parametrize("n,a", [(256, 6), (1, 2), (5000, 18), (33, 1)])
def test_injected_uniforms_bit_exact(n, a):
    from oracle.pg_loss import sample_categorical
    from rlpyt_b200.distributions.categorical import Categorical, DistInfo
    p = _probs(n, a, n)
    rng = np.random.default_rng(a)
    u = rng.random(n).astype(np.float32)
    edge = np.array([0.0, np.nextafter(np.float32(1), np.float32(0)), 0.5, 0.25], dtype=np.float32)
    u[:min(4, n)] = edge[:min(4, n)]
    want = sample_categorical(p, u)
    got = Categorical(dim=a).sample(DistInfo(prob=torch.from_numpy(p).cuda()), uniform=torch.from_numpy(u))
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), want)
    if a > 1:
        assert (np.take_along_axis(p, want[:, None], 1)[:, 0] > 0).mean() > 0.99      # never a zero-probability action (u < 1)


def test_philox_stream_counter_and_graph_replay():
    from oracle.philox import uniforms
    from oracle.pg_loss import sample_categorical
    from rlpyt_b200 import _lib
    from rlpyt_b200.distributions.categorical import Categorical, DistInfo
    n, a, seed = 512, 6, 123456789012345
    p = _probs(n, a, 1)
    pc = torch.from_numpy(p).cuda()
    state = torch.tensor([seed, 40], dtype=torch.int64, device="cuda")
    act = torch.empty(n, dtype=torch.int64, device="cuda")
    uo = torch.empty(n, dtype=torch.float32, device="cuda")
    for call in (40, 41, 42):
        _lib.call("rl_categorical_sample_f32", _lib.ptr(pc), None, _lib.ptr(state), _lib.ptr(act), _lib.ptr(uo), n, a, _lib.stream())
        u = uniforms(n, call, seed)
        assert np.array_equal(uo.cpu().numpy(), u)
        assert np.array_equal(act.cpu().numpy(), sample_categorical(p, u))
        assert int(state[1]) == call + 1
    # the distribution object: seeded stream, and a captured sample draws new numbers at every replay
    d = Categorical(dim=a)
    d.manual_seed(7, device="cuda")
    first = d.sample(DistInfo(prob=pc)).clone()
    assert np.array_equal(first.cpu().numpy(), sample_categorical(p, uniforms(n, 0, 7)))
    graph = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        drawn = d.sample(DistInfo(prob=pc))
    seen = []
    for _ in range(3):
        graph.replay()
        torch.cuda.synchronize()
        seen.append(drawn.clone())
    base = int(d._rng[1]) - 3
    for k, s in enumerate(seen):
        assert np.array_equal(s.cpu().numpy(), sample_categorical(p, uniforms(n, base + k, 7)))
    assert not torch.equal(seen[0], seen[1])


def test_frequencies_follow_probabilities():
    from rlpyt_b200.distributions.categorical import Categorical, DistInfo
    a, n = 6, 200_000
    row = torch.tensor([0.05, 0.3, 0.0, 0.15, 0.4, 0.1], device="cuda")
    d = Categorical(dim=a)
    d.manual_seed(11, device="cuda")
    draws = d.sample(DistInfo(prob=row.expand(n, a).contiguous()))
    freq = torch.bincount(draws, minlength=a).double().cpu().numpy() / n
    np.testing.assert_allclose(freq, row.cpu().numpy(), atol=5e-3)
    assert freq[2] == 0.0


@pytest.mark.parametrize("B,F,A", [(256, 512, 6), (128, 512, 18), (3, 64, 2), (1000, 512, 4)])
def test_fused_policy_head_vs_torch(B, F, A):
    """pg_head_sample_kernel = softmax(h W^T + b), h w_v + b_v and the inverse-CDF draw in one launch: probabilities
    and values against torch fp64 (1e-6), actions bit-exact against the oracle rule applied to the kernel's own
    probabilities and the oracle's Philox uniforms; the call counter advances once per launch."""
    from oracle.philox import uniforms
    from oracle.pg_loss import sample_categorical
    from rlpyt_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(B + A)
    h = torch.relu(torch.randn(B, F, device="cuda", generator=g))
    w_pi = torch.randn(A, F, device="cuda", generator=g) / F ** 0.5
    b_pi = torch.randn(A, device="cuda", generator=g) * 0.1
    w_v = torch.randn(F, device="cuda", generator=g) / F ** 0.5
    b_v = torch.randn(1, device="cuda", generator=g)
    state = torch.tensor([99, 5, 0], dtype=torch.int64, device="cuda")
    prob = torch.empty(B, A, device="cuda")
    val = torch.empty(B, device="cuda")
    act = torch.empty(B, dtype=torch.int64, device="cuda")
    for call in (5, 6):
        _lib.call("rl_pg_head_sample_f32", _lib.ptr(h), _lib.ptr(w_pi), _lib.ptr(b_pi), _lib.ptr(w_v), _lib.ptr(b_v), None,
                  _lib.ptr(state), _lib.ptr(prob), _lib.ptr(val), _lib.ptr(act), B, F, A, _lib.stream())
        ref_p = torch.softmax(h.double() @ w_pi.double().t() + b_pi.double(), -1)
        ref_v = h.double() @ w_v.double() + b_v.double()
        np.testing.assert_allclose(prob.cpu().numpy(), ref_p.cpu().numpy(), rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(val.cpu().numpy(), ref_v.cpu().numpy(), rtol=2e-6, atol=2e-6)
        assert np.array_equal(act.cpu().numpy(), sample_categorical(prob.cpu().numpy(), uniforms(B, call, 99)))
        assert state.tolist() == [99, call + 1, 0]


def test_agent_step_uses_the_fused_head_and_matches_the_training_forward():
    """AtariFfAgent.step on CUDA frames: recorded prob / value equal the differentiable forward (cuBLAS heads) to 1e-5,
    and the action is the oracle rule on the recorded probabilities with the distribution's Philox stream."""
    from collections import namedtuple
    from oracle.philox import uniforms
    from oracle.pg_loss import sample_categorical
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    Spaces = namedtuple("Spaces", "observation action")
    agent = AtariFfAgent()
    agent.initialize(Spaces(namedtuple("O", "shape")((4, 84, 84)), namedtuple("Ac", "n")(6)))
    agent.to_device(0)
    agent.distribution.manual_seed(21, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    obs = torch.randint(0, 256, (256, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    pa = torch.zeros(256, dtype=torch.int64, device="cuda")
    pr = torch.zeros(256, device="cuda")
    for call in range(2):
        step = agent.step(obs, pa, pr)
        with torch.no_grad():
            pi, v = agent.model(obs, None, None)
        np.testing.assert_allclose(step.agent_info.dist_info.prob.cpu().numpy(), pi.cpu().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(step.agent_info.value.cpu().numpy(), v.cpu().numpy(), rtol=1e-5, atol=1e-5)
        want = sample_categorical(step.agent_info.dist_info.prob.cpu().numpy(), uniforms(256, call, 21))
        assert np.array_equal(step.action.cpu().numpy(), want)
