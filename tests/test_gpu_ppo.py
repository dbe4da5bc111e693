"""GPU parity of the algorithm layer (PPO / A2C optimize_agent on device-resident samples, the
row-gather kernels and the fused clip+Adam step) against the reference's recorded outputs
(tests/golden/ppo.npz) and torch's own clip_grad_norm_/Adam."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from collections import namedtuple  # noqa: E402

T, B, IMAGE, A = 8, 6, (4, 36, 36), 5
Spaces = namedtuple("Spaces", "observation action")
Obs = namedtuple("Obs", "shape")
Act = namedtuple("Act", "n")


def make_agent(sd0):
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    agent = AtariFfAgent(initial_model_state_dict={k: v.clone() for k, v in sd0.items()})
    agent.initialize(Spaces(Obs(IMAGE), Act(A)))
    agent.to_device(0)
    return agent


def make_samples(g, name, itr, device="cuda"):
    from rlpyt_b200.samplers.collections import Samples, AgentSamplesBsv, EnvSamples
    from rlpyt_b200.agents.pg.base import AgentInfo
    from rlpyt_b200.distributions.categorical import DistInfo
    t = lambda k: torch.from_numpy(g[f"{name}/itr{itr}/{k}"]).to(device)
    all_action = torch.cat([torch.zeros(1, B, dtype=torch.int64, device=device), t("action")])
    all_reward = torch.cat([torch.zeros(1, B, device=device), t("reward")])
    return Samples(
        agent=AgentSamplesBsv(action=all_action[1:], prev_action=all_action[:-1],
                              agent_info=AgentInfo(dist_info=DistInfo(prob=t("old_prob")), value=t("value")),
                              bootstrap_value=t("bv")),
        env=EnvSamples(observation=t("obs"), reward=all_reward[1:], prev_reward=all_reward[:-1],
                       done=t("done"), env_info=None))


CFG = {
    "ppo": (dict(gae_lambda=0.98, minibatches=2, epochs=2), True),
    "ppo_valid_norm": (dict(gae_lambda=1, minibatches=3, epochs=2, normalize_advantage=True, ratio_clip=0.2), False),
}


@pytest.mark.parametrize("name", list(CFG))
@pytest.mark.parametrize("samples_on", ["cuda", "cpu"])
def test_ppo_two_iterations_vs_reference(golden, name, samples_on):
    """Same weights, same samples, same numpy shuffle stream -> same OptInfo rows and weights as the
    reference's CPU run.  First update within 1e-5 (north_star); later rows inherit fp32
    conv/GEMM summation-order differences through Adam, held to 2e-4."""
    from oracle import atari_ff
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.samplers.collections import BatchSpec
    g = golden("ppo")
    sd0 = atari_ff.init_state_dict(IMAGE, A, seed=int(g[f"{name}/sd0_seed"][0]))
    agent = make_agent(sd0)
    kwargs, mbr = CFG[name]
    algo = PPO(**kwargs)
    algo.initialize(agent, 4, BatchSpec(T, B), mid_batch_reset=mbr)
    np.random.seed(77)
    for itr in range(2):
        agent.train_mode(itr)
        info = algo.optimize_agent(itr, make_samples(g, name, itr, samples_on))
        for f in ("loss", "gradNorm", "entropy", "perplexity"):
            got, want = np.asarray(getattr(info, f)), g[f"{name}/itr{itr}/opt_{f}"]
            assert got.shape == want.shape
            if itr == 0:
                np.testing.assert_allclose(got[0], want[0], rtol=1e-5, atol=1e-7, err_msg=f)
            np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-6, err_msg=f)
        sd = agent.state_dict()
        for k, v in sd.items():
            want = g[f"{name}/itr{itr}/sd/{k}"]
            got = v.cpu().numpy()[:8] if k == "conv.head.model.0.weight" else v.cpu().numpy()
            # Adam moves an element whose gradient is at fp32-noise level by up to ~lr/2 per step
            # whichever way the noise points; everything else agrees to 1e-3 relative
            np.testing.assert_allclose(got, want, rtol=1e-3, atol=5e-4, err_msg=k)


@pytest.mark.parametrize("image", [(4, 84, 84), (4, 104, 80)])
def test_ppo_first_update_full_config_vs_oracle(image):
    """BASELINE.json configs[2] at FULL size - T=128, B=256, 8192-sample minibatches, the persistent conv kernels
    with 55 frames per CTA, the 8192 x 512 x 3200 GEMMs, the row-gather over a 925 MB batch - not only the
    [T=8,B=6,(4,36,36)] case of the golden file: the first minibatch update of ``PPO.optimize_agent`` against
    oracle/ppo.py (torch-CPU fp32, pinned bit for bit to the reference by tests/test_oracle_*.py) on identical
    samples, weights and numpy shuffle stream.  North-star tolerance: 1e-5 relative, with float64 as the arbiter where
    the reference's own fp32 summation order is further than that from exact arithmetic (see the loop below)."""
    from oracle import atari_ff
    from oracle.ppo import PpoOracle
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    from rlpyt_b200.agents.pg.base import AgentInfo
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.distributions.categorical import DistInfo
    from rlpyt_b200.samplers.collections import AgentSamplesBsv, BatchSpec, EnvSamples, Samples
    Tf, Bf, Af = 128, 256, 6
    rng = np.random.default_rng(5)
    obs = rng.integers(0, 256, size=(Tf, Bf) + image, dtype=np.uint8)
    action = rng.integers(0, Af, size=(Tf + 1, Bf))
    reward = rng.choice(np.array([-1, 0, 1], np.float32), size=(Tf + 1, Bf), p=[.02, .96, .02]).astype(np.float32)
    done = rng.random((Tf, Bf)) < 1 / 500.
    value = rng.standard_normal((Tf, Bf)).astype(np.float32)
    prob = rng.dirichlet(np.ones(Af), (Tf, Bf)).astype(np.float32)
    bv = rng.standard_normal((1, Bf)).astype(np.float32)
    sd0 = atari_ff.init_state_dict(image, Af, seed=3)
    kw = dict(discount=0.99, learning_rate=1e-3, value_loss_coeff=1., entropy_loss_coeff=0.01, clip_grad_norm=1.,
              gae_lambda=0.98, linear_lr_schedule=True, minibatches=4, epochs=4, ratio_clip=0.1)
    oracle = PpoOracle(sd0, n_itr=100, **kw)
    np.random.seed(123)
    want = oracle.optimize_agent(0, obs, action[1:], reward[1:], done, value, prob, bv, max_updates=1)
    # the same update in float64: where two fp32 implementations differ only by summation order (8192 x 400 positions
    # per weight-gradient element, heavy cancellation) this is the arbiter
    oracle64 = PpoOracle({k: v.double() for k, v in sd0.items()}, n_itr=100, **kw)
    np.random.seed(123)
    exact = oracle64.optimize_agent(0, obs, action[1:], reward[1:], done, value.astype(np.float64), prob.astype(np.float64),
                                    bv, max_updates=1)
    agent = AtariFfAgent(initial_model_state_dict={k: v.clone() for k, v in sd0.items()})
    agent.initialize(Spaces(Obs(image), Act(Af)))
    agent.to_device(0)
    algo = PPO(**kw)
    algo.initialize(agent, 100, BatchSpec(Tf, Bf), mid_batch_reset=True)
    cu = lambda a: torch.from_numpy(a).cuda()
    all_a, all_r = cu(action), cu(reward)
    samples = Samples(
        agent=AgentSamplesBsv(action=all_a[1:], prev_action=all_a[:-1],
                              agent_info=AgentInfo(dist_info=DistInfo(prob=cu(prob)), value=cu(value)), bootstrap_value=cu(bv)),
        env=EnvSamples(observation=cu(obs), reward=all_r[1:], prev_reward=all_r[:-1], done=cu(done), env_info=None))
    np.random.seed(123)
    agent.train_mode(0)
    info = algo.optimize_agent(0, samples)
    assert len(info.loss) == 16 and all(np.isfinite(info.loss)) and all(np.isfinite(info.gradNorm))
    for f in ("loss", "gradNorm", "entropy", "perplexity"):
        got, ref32, ref64 = float(np.asarray(getattr(info, f))[0]), float(want[f][0]), float(exact[f][0])
        # within 1e-5 of the reference arithmetic (north_star) - or, where the reference's own fp32 summation is further
        # than that from exact arithmetic (the gradient norm at this size), at least as close to exact as it is
        tol = max(1e-5 * abs(ref32) + 1e-7, 1.5 * abs(ref32 - ref64))
        assert abs(got - ref32) <= tol or abs(got - ref64) <= abs(ref32 - ref64), (f, got, ref32, ref64)
        assert abs(got - ref64) <= 1e-4 * abs(ref64) + 1e-6, (f, got, ref64)


def test_a2c_two_iterations_vs_reference(golden):
    from oracle import atari_ff
    from rlpyt_b200.algos.pg.a2c import A2C
    from rlpyt_b200.samplers.collections import BatchSpec
    g = golden("ppo")
    sd0 = atari_ff.init_state_dict(IMAGE, A, seed=int(g["a2c/sd0_seed"][0]))
    agent = make_agent(sd0)
    algo = A2C(gae_lambda=0.95)
    algo.initialize(agent, 4, BatchSpec(T, B), mid_batch_reset=True)
    for itr in range(2):
        agent.train_mode(itr)
        info = algo.optimize_agent(itr, make_samples(g, "a2c", itr))
        for f in ("loss", "gradNorm", "entropy", "perplexity"):
            np.testing.assert_allclose(getattr(info, f), g[f"a2c/itr{itr}/opt_{f}"][0], rtol=2e-4, atol=1e-6,
                                       err_msg=f)


@pytest.mark.parametrize("shape,dtype", [((1000, 4, 84, 84), torch.uint8), ((333, 7), torch.float32),
                                         ((50, 3, 5), torch.int64), ((64, 1, 1), torch.uint8),
                                         ((77, 10), torch.uint8)])
def test_gather_rows_bit_exact(shape, dtype):
    from rlpyt_b200.utils.gather import gather_rows
    gen = torch.Generator(device="cuda").manual_seed(1)
    if dtype.is_floating_point:
        src = torch.randn(shape, device="cuda", generator=gen)
    else:
        src = torch.randint(0, 200, shape, device="cuda", generator=gen).to(dtype)
    idx = torch.randint(0, shape[0], (517,), device="cuda", generator=gen)
    assert torch.equal(gather_rows(src, idx), src[idx])
    assert gather_rows(src, idx[:0]).shape[0] == 0


def test_gather_rows_multi_bit_exact():
    from rlpyt_b200.utils.gather import gather_rows_multi
    gen = torch.Generator(device="cuda").manual_seed(2)
    n = 32768
    srcs = [torch.randint(0, 6, (n,), device="cuda", generator=gen),            # int64 action
            torch.randn(n, device="cuda", generator=gen),                       # f32
            torch.randn(n, 6, device="cuda", generator=gen),                    # prob
            torch.randn(n, device="cuda", generator=gen)]
    idx = torch.randperm(n, device="cuda", generator=gen)[:8192]
    outs = gather_rows_multi(srcs, idx)
    for s, o in zip(srcs, outs):
        assert torch.equal(o, s[idx])


def test_flat_adam_matches_torch_clip_and_adam():
    from rlpyt_b200.algos.optim import FlatAdam
    torch.manual_seed(0)
    shapes = [(16, 4, 8, 8), (16,), (512, 100), (5, 512), (5,), (1, 3)]
    ref_params = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    my_params = [p.detach().clone().requires_grad_(True) for p in ref_params]
    ref = torch.optim.Adam(ref_params, lr=1e-3)
    mine = FlatAdam(my_params, lr=1e-3)
    mine.direct_grads = False          # this test feeds gradients by hand into the attached .grad views of the flat buffer
    for step in range(5):
        grads = [torch.randn(s, device="cuda") * (10.0 if step % 2 == 0 else 0.01) for s in shapes]
        mine.zero_grad()
        for p, q, gr in zip(ref_params, my_params, grads):
            p.grad = gr.clone()
            q.grad.add_(gr)  # accumulate into the persistent flat view, as autograd does
        want_norm = torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
        ref.step()
        norm = mine.clip_and_step(1.0)
        np.testing.assert_allclose(norm.item(), want_norm.item(), rtol=1e-6)
        for p, q in zip(ref_params, my_params):
            np.testing.assert_allclose(q.detach().cpu().numpy(), p.detach().cpu().numpy(), rtol=2e-6, atol=1e-7)
    # state_dict round trip in torch.optim.Adam's format
    sd = mine.state_dict()
    again = torch.optim.Adam([p.detach().clone().requires_grad_(True) for p in ref_params], lr=1e-3)
    again.load_state_dict(sd)
    assert int(again.state_dict()["state"][0]["step"]) == 5
    mine2 = FlatAdam([p.detach().clone().requires_grad_(True) for p in ref_params], lr=1e-3)
    mine2.load_state_dict(ref.state_dict())
    assert mine2.step_count == 5
    np.testing.assert_allclose(mine2.exp_avg.cpu().numpy(), mine.exp_avg.cpu().numpy(), rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize("shape", [(4, 84, 84), (4, 36, 36), (4, 104, 80)])
@pytest.mark.parametrize("use_rows", [False, True])
def test_fused_first_layer_matches_torch_conv(shape, use_rows):
    """csrc/conv1.cu (u8 gather + /255 + conv 8x8/s4 + bias + ReLU, and its weight/bias gradient)
    against torch's fp32 conv2d on the converted image: 1e-5 relative (summation order only)."""
    import torch.nn.functional as F
    from rlpyt_b200.models.conv1_op import conv1_u8_relu
    torch.backends.cudnn.allow_tf32 = False
    gen = torch.Generator(device="cuda").manual_seed(3)
    R = 300
    obs = torch.randint(0, 256, (R,) + shape, dtype=torch.uint8, device="cuda", generator=gen)
    rows = torch.randint(0, R, (77,), device="cuda", generator=gen) if use_rows else None
    w = (torch.rand(16, 4, 8, 8, device="cuda", generator=gen) - 0.5) / 8
    b = (torch.rand(16, device="cuda", generator=gen) - 0.5) / 8
    w1, b1 = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    w2, b2 = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = conv1_u8_relu(w1, b1, obs, rows)
    x = (obs if rows is None else obs[rows]).float().mul_(1. / 255)
    y_ref = F.relu(F.conv2d(x, w2, b2, stride=4))
    assert y.shape == y_ref.shape
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_ref.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)
    g = torch.randn(y.shape, device="cuda", generator=gen)
    y.backward(g)
    # Reference gradient with the SAME ReLU mask (y > 0 of the fused output): a pre-activation within
    # fp32 noise of zero may flip sign between two summation orders, which would move 256 taps by O(1).
    gm = g * (y.detach() > 0)
    gw_ref = torch.nn.grad.conv2d_weight(x, w.shape, gm, stride=4)
    gb_ref = gm.sum((0, 2, 3))
    scale = float(gw_ref.abs().max())
    np.testing.assert_allclose(w1.grad.cpu().numpy(), gw_ref.cpu().numpy(), rtol=1e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(b1.grad.cpu().numpy(), gb_ref.cpu().numpy(), rtol=1e-4, atol=2e-5 * float(gb_ref.abs().max()))
    # and the masks themselves differ on at most a handful of noise-level elements
    assert int(((y.detach() > 0) != (y_ref.detach() > 0)).sum()) <= 8


def test_model_fused_and_generic_paths_agree():
    from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
    torch.manual_seed(0)
    m = AtariFfModel((4, 84, 84), 6).cuda()
    assert m.fused_first_layer
    obs = torch.randint(0, 256, (3, 5, 4, 84, 84), dtype=torch.uint8, device="cuda")
    pi, v = m(obs, None, None)
    m.fused_first_layer = False
    pi2, v2 = m(obs, None, None)
    assert pi.shape == (3, 5, 6) and v.shape == (3, 5)
    np.testing.assert_allclose(pi.detach().cpu().numpy(), pi2.detach().cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(v.detach().cpu().numpy(), v2.detach().cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_model_tensor_core_head_matches_cublas_head():
    """Minibatch-sized input takes the tcgen05 3xTF32 fc layer; outputs and all gradients agree with
    the cuBLAS fp32 path to 1e-5 / 1e-4."""
    from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
    torch.manual_seed(1)
    m = AtariFfModel((4, 84, 84), 6).cuda()
    obs = torch.randint(0, 256, (2048, 4, 84, 84), dtype=torch.uint8, device="cuda")
    w = torch.randn(2048, device="cuda")

    def run(min_rows):
        m.TC_GEMM_MIN_ROWS = min_rows
        m.zero_grad()
        pi, v = m(obs, None, None)
        ((pi[:, 0] * w).sum() + (v * w).sum()).backward()
        return pi.detach().clone(), v.detach().clone(), [p.grad.detach().clone() for p in m.parameters()]
    pi1, v1, g1 = run(2048)       # tensor-core head
    pi2, v2, g2 = run(1 << 30)    # cuBLAS head
    np.testing.assert_allclose(pi1.cpu().numpy(), pi2.cpu().numpy(), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(v1.cpu().numpy(), v2.cpu().numpy(), rtol=1e-5, atol=2e-6)
    for a, b in zip(g1, g2):
        s = float(b.abs().max()) + 1e-12
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-3, atol=2e-5 * s)


@pytest.mark.parametrize("normalize,image", [(False, (4, 84, 84)), (True, (4, 36, 36))])
def test_minibatch_cuda_graph_is_bit_identical_to_eager_issue(normalize, image, monkeypatch):
    """PPO.optimize_agent replays the minibatch body (gather -> forward -> fused loss -> backward) as ONE CUDA graph from
    the second iteration on (algos/pg/ppo.py ``_minibatch_graph``).  Same kernels, same order, same inputs: four
    iterations over a sample buffer that is rewritten IN PLACE between iterations (what the samplers do) must leave
    parameters, Adam state and OptInfo bit-identical to the eagerly issued run - including the linearly scheduled
    learning rate and ratio clip (the clip reaches the captured loss kernel through a device scalar)."""
    from oracle import atari_ff
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    from rlpyt_b200.agents.pg.base import AgentInfo
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.distributions.categorical import DistInfo
    from rlpyt_b200.samplers.collections import AgentSamplesBsv, BatchSpec, EnvSamples, Samples
    Tn, Bn, An = 16, 32, 6
    sd0 = atari_ff.init_state_dict(image, An, seed=7)

    def run(graph):
        monkeypatch.setenv("RLPYT_B200_LEARNER_GRAPH", "1" if graph else "0")
        agent = AtariFfAgent(initial_model_state_dict={k: v.clone() for k, v in sd0.items()})
        agent.initialize(Spaces(Obs(image), Act(An)))
        agent.to_device(0)
        algo = PPO(gae_lambda=0.95, minibatches=4, epochs=2, normalize_advantage=normalize, ratio_clip=0.2)
        algo.initialize(agent, 10, BatchSpec(Tn, Bn), mid_batch_reset=not normalize)
        g = torch.Generator(device="cuda").manual_seed(11)
        all_a = torch.zeros(Tn + 1, Bn, dtype=torch.int64, device="cuda")
        all_r = torch.zeros(Tn + 1, Bn, device="cuda")
        samples = Samples(
            agent=AgentSamplesBsv(action=all_a[1:], prev_action=all_a[:-1],
                                  agent_info=AgentInfo(dist_info=DistInfo(prob=torch.zeros(Tn, Bn, An, device="cuda")),
                                                       value=torch.zeros(Tn, Bn, device="cuda")),
                                  bootstrap_value=torch.zeros(1, Bn, device="cuda")),
            env=EnvSamples(observation=torch.zeros((Tn, Bn) + image, dtype=torch.uint8, device="cuda"), reward=all_r[1:],
                           prev_reward=all_r[:-1], done=torch.zeros(Tn, Bn, dtype=torch.bool, device="cuda"), env_info=None))
        infos = []
        np.random.seed(9)
        agent.train_mode(0)
        for itr in range(4):
            samples.env.observation.copy_(torch.randint(0, 256, samples.env.observation.shape, dtype=torch.uint8, device="cuda", generator=g))
            all_a.copy_(torch.randint(0, An, all_a.shape, device="cuda", generator=g))
            all_r.copy_(torch.randn(all_r.shape, device="cuda", generator=g))
            samples.env.done.copy_(torch.rand(Tn, Bn, device="cuda", generator=g) < 0.05)
            samples.agent.agent_info.value.copy_(torch.randn(Tn, Bn, device="cuda", generator=g))
            samples.agent.agent_info.dist_info.prob.copy_(torch.softmax(torch.randn(Tn, Bn, An, device="cuda", generator=g), -1))
            samples.agent.bootstrap_value.copy_(torch.randn(1, Bn, device="cuda", generator=g))
            infos.append(algo.optimize_agent(itr, samples))
        n_graphs = len(algo.__dict__.get("_mb_graphs") or {})
        return algo.optimizer.flat_param.clone(), algo.optimizer.exp_avg_sq.clone(), infos, n_graphs

    p_e, v_e, i_e, g_e = run(False)
    p_g, v_g, i_g, g_g = run(True)
    assert g_e == 0 and g_g == 1                                   # one capture served iterations 1..3
    assert torch.equal(p_e, p_g) and torch.equal(v_e, v_g)
    for a, b in zip(i_e, i_g):
        for f in ("loss", "gradNorm", "entropy", "perplexity"):
            assert getattr(a, f) == getattr(b, f), f



def test_relu_backward_fused_into_fc_input_gradient_is_bit_identical(monkeypatch):
    """The second conv layer's ReLU backward rides in the epilogue of the fc layer's input-gradient GEMM
    (rl_gemm_ts_masked_f32) at minibatch sizes: every parameter gradient equals, bit for bit, the one computed with
    the separate ReLU-backward pass."""
    from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
    torch.manual_seed(3)
    model = AtariFfModel((4, 84, 84), 6).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randint(0, 256, (1536, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    gp = torch.randn(1536, 6, device="cuda", generator=g)
    gv = torch.randn(1536, device="cuda", generator=g)

    def grads(fused):
        monkeypatch.setenv("RLPYT_B200_FUSE_RELU_BWD", "1" if fused else "0")
        model.zero_grad(set_to_none=True)
        pi, v = model(obs, None, None)
        torch.autograd.backward([pi, v], [gp, gv])
        return {k: p.grad.clone() for k, p in model.named_parameters()}

    a, b = grads(False), grads(True)
    assert all(float(x.abs().max()) > 0 for x in a.values())
    for k in a:
        assert torch.equal(a[k], b[k]), k



def test_channel_absmax_from_conv2_dgrad_epilogue_is_bit_identical(monkeypatch):
    """conv2's input-gradient kernel leaves max |grad[:, c]| per channel in device memory (rl_conv2_dgrad_s2d_absmax) and the
    first layer's kind::i8 weight gradient takes it from there (rl_conv1_u8_wgrad_i8_scaled) instead of running its own
    absmax pass: the maxima equal torch's, and every parameter gradient of the model is bit-identical with and without
    the hand-over."""
    import importlib
    from rlpyt_b200.models import conv2_op
    from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
    torch.manual_seed(5)
    model = AtariFfModel((4, 84, 84), 6).cuda()
    g = torch.Generator(device="cuda").manual_seed(2)
    obs = torch.randint(0, 256, (1100, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    gp = torch.randn(1100, 6, device="cuda", generator=g)
    gv = torch.randn(1100, device="cuda", generator=g)
    seen = {}

    def grads(fused):
        monkeypatch.setattr(conv2_op, "FUSE_ABSMAX", fused)
        conv2_op.LAST_DGRAD_ABSMAX.clear()
        model.zero_grad(set_to_none=True)
        pi, v = model(obs, None, None)
        if fused:      # look at what the epilogue produced before the first layer consumes it
            from rlpyt_b200.models import conv1_op
            orig = conv1_op._producer_absmax

            def spy(gr):
                out = orig(gr)
                seen["absmax"], seen["ref"] = out, gr.abs().amax(dim=(0, 2, 3))
                return out
            monkeypatch.setattr(conv1_op, "_producer_absmax", spy)
        torch.autograd.backward([pi, v], [gp, gv])
        return {k: p.grad.clone() for k, p in model.named_parameters()}

    a, b = grads(False), grads(True)
    assert seen["absmax"] is not None and torch.equal(seen["absmax"], seen["ref"])
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_direct_gradient_writes_into_the_flat_buffer(monkeypatch):
    """With ``FlatAdam`` the hand-written backward kernels write each parameter gradient straight into its slot of the flat
    gradient buffer (algos/optim.py ``grad_destination``) and autograd adopts the slot view as ``.grad`` - no add kernel per
    parameter.  Checked: (i) every ``.grad`` of the AtariFf network IS its slot after backward; (ii) the flat buffer equals,
    bit for bit, the one produced by accumulating into attached ``.grad`` views; (iii) a network used TWICE in one graph
    still accumulates both contributions; (iv) a gradient that arrives some other way is collected into its slot."""
    from rlpyt_b200.algos.optim import FlatAdam
    from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
    g = torch.Generator(device="cuda").manual_seed(4)
    obs = torch.randint(0, 256, (1200, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    obs2 = torch.randint(0, 256, (1200, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    gp, gv = torch.randn(1200, 6, device="cuda", generator=g), torch.randn(1200, device="cuda", generator=g)

    def run(direct, twice):
        torch.manual_seed(9)
        model = AtariFfModel((4, 84, 84), 6).cuda()
        opt = FlatAdam(model.parameters(), lr=1e-3)
        opt.direct_grads = direct
        opt.zero_grad()
        pi, v = model(obs, None, None)
        if twice:
            pi2, v2 = model(obs2, None, None)
            pi, v = pi + pi2, v + v2
        torch.autograd.backward([pi, v], [gp, gv])
        base = opt.flat_grad.data_ptr()
        in_slot = [p.grad is not None and p.grad.data_ptr() == base + 4 * off for p, off in zip(opt.param_groups[0]["params"], opt._offsets)]
        if direct:
            opt._collect_grads()
        return opt.flat_grad.clone(), in_slot, opt, model

    ref, _, _, _ = run(False, False)
    got, in_slot, opt, model = run(True, False)
    assert all(in_slot)                                              # (i): adopted, not copied
    assert torch.equal(ref, got)                                     # (ii)
    ref2, _, _, _ = run(False, True)
    got2, in_slot2, _, _ = run(True, True)
    assert torch.allclose(ref2, got2, rtol=1e-6, atol=1e-7) and float(ref2.abs().max()) > 0      # (iii): a + b vs b + a per element
    # (iv) a foreign gradient
    opt.zero_grad()
    p0 = opt.param_groups[0]["params"][0]
    p0.grad = torch.full_like(p0, 3.0)
    opt._collect_grads()
    off = opt._offsets[0]
    assert float(opt.flat_grad[off:off + p0.numel()].min()) == 3.0
    p0.grad.fill_(5.0)                                               # the same foreign buffer rewritten (a replayed CUDA graph does that)
    opt._collect_grads()
    assert float(opt.flat_grad[off:off + p0.numel()].min()) == 5.0
