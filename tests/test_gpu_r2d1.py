"""R2D1 on the device (rlpyt_b200.algos.dqn.r2d1 + AtariR2d1Agent / AtariR2d1Model) against the reference run on the CPU
(tests/golden/r2d1.npz, tests/golden/make_golden.py: gen_r2d1): same weights, same sampled batch -> loss, valid TD errors,
sequence priorities and the gradient of every parameter; input priorities of a sampler batch; the value rescaling."""
from collections import namedtuple

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = ["r2d1_double_pri", "r2d1_plain_uniform_huber", "r2d1_dueling"]
IMAGE_SHAPE = (4, 36, 36)


def build(g, name):
    from rlpyt_b200.agents.dqn.atari.atari_r2d1_agent import AtariR2d1Agent
    from rlpyt_b200.algos.dqn.r2d1 import R2D1
    (seed, wT, bT, n, B, double, prioritized, delta_clip, rsi, dueling, A, H) = g[f"{name}/cfg"]
    seed, wT, bT, n, B, rsi, A, H = (int(v) for v in (seed, wT, bT, n, B, rsi, A, H))
    Spaces = namedtuple("Spaces", "observation action")
    agent = AtariR2d1Agent(model_kwargs=dict(channels=[4, 8, 8], fc_size=32, lstm_size=H, head_size=16, dueling=bool(dueling)))
    agent.initialize(Spaces(namedtuple("O", "shape")(IMAGE_SHAPE), namedtuple("Ac", "n")(A)))
    agent.model.load_state_dict({k[len(name) + 7:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}/model/")})
    agent.target_model.load_state_dict({k[len(name) + 8:]: torch.from_numpy(g[k]) for k in g.files
                                        if k.startswith(f"{name}/target/")})
    agent.to_device(0)
    algo = R2D1(discount=0.99, batch_T=bT, batch_B=B, warmup_T=wT, store_rnn_state_interval=rsi, n_step_return=n,
                double_dqn=bool(double), prioritized_replay=bool(prioritized), delta_clip=None if delta_clip < 0 else float(delta_clip),
                pri_eta=0.9, input_priority_shift=0 if rsi == 0 else None)
    algo.agent = agent
    return agent, algo, dict(seed=seed, wT=wT, bT=bT, n=n, B=B, rsi=rsi, prioritized=bool(prioritized))


@pytest.mark.parametrize("name", CASES)
def test_r2d1_loss_priorities_and_gradients_vs_reference(golden, name):
    from rlpyt_b200.models.dqn.atari_r2d1_model import RnnState
    from rlpyt_b200.replays.sequence.n_step import SamplesFromReplay
    from rlpyt_b200.replays.sequence.prioritized import SamplesFromReplayPri
    g = golden("r2d1")
    agent, algo, c = build(g, name)
    L = c["wT"] + c["bT"] + c["n"]
    obs = np.random.default_rng(c["seed"]).integers(0, 256, size=(L, c["B"]) + IMAGE_SHAPE, dtype=np.uint8)
    assert int(obs.astype(np.int64).sum()) == int(g[f"{name}/batch/all_observation_sum"][0])
    cu = lambda k: torch.from_numpy(g[f"{name}/batch/{k}"]).cuda()
    init = None if c["rsi"] == 0 else RnnState(h=cu("init_h"), c=cu("init_c"))
    base = SamplesFromReplay(all_observation=torch.from_numpy(obs).cuda(), all_action=cu("all_action"), all_reward=cu("all_reward"),
                             return_=cu("return_"), done=cu("done"), done_n=cu("done_n"), init_rnn_state=init)
    samples = SamplesFromReplayPri(*base, is_weights=cu("is_weights")) if c["prioritized"] else base
    agent.train_mode(0)
    loss, td, pri = algo.loss(samples)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"{name}/loss"][0], rtol=2e-5)
    want_td = g[f"{name}/td_abs_errors"]
    np.testing.assert_allclose(td.cpu().numpy(), want_td, rtol=1e-4, atol=2e-5 * np.abs(want_td).max())
    np.testing.assert_allclose(pri.detach().cpu().numpy(), g[f"{name}/priorities"], rtol=1e-4)
    assert np.array_equal(td.cpu().numpy() == 0, want_td == 0)                  # the same steps are masked out
    # absolute floor from the largest gradient of the network: the dueling head's shared advantage bias has an exactly
    # zero gradient in real arithmetic (it cancels in A - mean(A)); what both sides hold there is rounding noise
    floor = 1e-6 * max(float(np.abs(g[f"{name}/grad/{k}"]).max()) for k, _ in agent.model.named_parameters())
    for k, p in agent.model.named_parameters():
        want = g[f"{name}/grad/{k}"]
        got = (p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=1e-3, atol=max(floor, 2e-5 * np.abs(want).max()), err_msg=k)


@pytest.mark.parametrize("name", ["r2d1_double_pri", "r2d1_dueling"])
def test_r2d1_input_priorities_vs_reference(golden, name):
    from rlpyt_b200.agents.dqn.r2d1_agent import AgentInfo
    from rlpyt_b200.samplers.collections import AgentSamples, EnvSamples, Samples
    g = golden("r2d1")
    _, algo, _ = build(g, name)
    cu = lambda k: torch.from_numpy(g[f"{name}/input/{k}"]).cuda()
    smp = Samples(agent=AgentSamples(action=cu("action"), prev_action=cu("action"), agent_info=AgentInfo(q=cu("q"), prev_rnn_state=None)),
                  env=EnvSamples(observation=None, reward=cu("reward"), prev_reward=cu("reward"), done=cu("done"), env_info=None))
    pri = algo.compute_input_priorities(smp)
    assert pri.is_cuda
    np.testing.assert_allclose(pri.cpu().numpy(), g[f"{name}/input/priorities"], rtol=2e-5)


def test_value_rescaling_vs_reference(golden):
    from rlpyt_b200.algos.dqn.r2d1 import R2D1
    g = golden("r2d1")
    algo = R2D1()
    x = torch.from_numpy(g["value_scale/x"]).cuda()
    np.testing.assert_allclose(algo.value_scale(x).cpu().numpy(), g["value_scale/h"], rtol=1e-6)
    np.testing.assert_allclose(algo.inv_value_scale(x).cpu().numpy(), g["value_scale/h_inv"], rtol=2e-5)


def test_r2d1_end_to_end_learns_from_sequence_replay():
    """sampler-shaped [T,B] batches -> samples_to_buffer (stored RNN state + input priorities) -> prioritized sequence
    replay in HBM -> updates; the public call sequence of the reference runner for this algorithm."""
    from rlpyt_b200.agents.dqn.atari.atari_r2d1_agent import AtariR2d1Agent
    from rlpyt_b200.agents.dqn.r2d1_agent import AgentInfo
    from rlpyt_b200.algos.dqn.r2d1 import R2D1
    from rlpyt_b200.models.dqn.atari_r2d1_model import RnnState
    from rlpyt_b200.samplers.collections import AgentSamples, BatchSpec, EnvSamples, Samples
    A, H, T, B = 4, 16, 8, 6
    Spaces = namedtuple("Spaces", "observation action")
    agent = AtariR2d1Agent(model_kwargs=dict(channels=[4, 8, 8], fc_size=32, lstm_size=H, head_size=16))
    agent.initialize(Spaces(namedtuple("O", "shape")(IMAGE_SHAPE), namedtuple("Ac", "n")(A)))
    agent.to_device(0)
    algo = R2D1(batch_T=8, batch_B=4, warmup_T=8, store_rnn_state_interval=8, n_step_return=3, min_steps_learn=T * B * 5,
                replay_size=T * B * 16, replay_ratio=1, target_update_interval=2)
    examples = dict(observation=np.zeros(IMAGE_SHAPE, np.uint8), action=np.int64(0), reward=np.float32(0), done=np.bool_(False),
                    agent_info=AgentInfo(q=np.zeros(A, np.float32),
                                         prev_rnn_state=RnnState(h=np.zeros((1, H), np.float32), c=np.zeros((1, H), np.float32))))
    algo.initialize(agent, 20, BatchSpec(T, B), mid_batch_reset=False, examples=examples)
    g = torch.Generator(device="cuda").manual_seed(0)
    losses = []
    for itr in range(12):
        agent.sample_mode(itr)
        agent.reset()
        obs = torch.randint(0, 256, (T, B) + IMAGE_SHAPE, dtype=torch.uint8, device="cuda", generator=g)
        acts, qs, hs, cs = [], [], [], []
        pa, pr = torch.zeros(B, dtype=torch.int64, device="cuda"), torch.zeros(B, device="cuda")
        for t in range(T):
            step = agent.step(obs[t], pa, pr)
            acts.append(step.action); qs.append(step.agent_info.q)
            hs.append(step.agent_info.prev_rnn_state.h); cs.append(step.agent_info.prev_rnn_state.c)
            pa, pr = step.action, torch.randn(B, device="cuda", generator=g)
        action = torch.stack(acts)
        samples = Samples(agent=AgentSamples(action=action, prev_action=action,
                                             agent_info=AgentInfo(q=torch.stack(qs), prev_rnn_state=RnnState(h=torch.stack(hs), c=torch.stack(cs)))),
                          env=EnvSamples(observation=obs, reward=torch.randn(T, B, device="cuda", generator=g), prev_reward=None,
                                         done=torch.rand(T, B, device="cuda", generator=g) < 0.03, env_info=None))
        assert samples.agent.agent_info.prev_rnn_state.h.shape == (T, B, 1, H)
        agent.train_mode(itr)
        info = algo.optimize_agent(itr, samples)
        losses.extend(info.loss)
        if itr >= 5:
            assert len(info.loss) == algo.updates_per_optimize and len(info.priority) == 4 * len(info.loss)
    assert len(losses) >= 7 and all(np.isfinite(losses))
    assert float(algo.replay_buffer.priority_tree.tree[0]) > 0
