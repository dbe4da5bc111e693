"""SURVEY.md 8(f) row 3, recurrent half: AtariLstmAgent (conv trunk on this repo's kernels + LSTM) with the recurrent
branches of PPO / A2C against the reference's recorded CPU run (tests/golden/ppo_lstm.npz: reference AtariLstmAgent +
PPO(minibatches over B, whole trajectories, valid mask) / A2C), and the recurrent-state handling of the samplers."""
from collections import namedtuple

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

T, B, IMAGE, A, H = 8, 6, (4, 36, 36), 5, 64
Spaces = namedtuple("Spaces", "observation action")
Obs = namedtuple("Obs", "shape")
Act = namedtuple("Act", "n")


def _agent(g, name):
    from rlpyt_b200.agents.pg.atari import AtariLstmAgent
    sd0 = {k[len(name) + 5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(f"{name}/sd0/")}
    agent = AtariLstmAgent(model_kwargs=dict(fc_sizes=128, lstm_size=H), initial_model_state_dict=sd0)
    agent.initialize(Spaces(Obs(IMAGE), Act(A)))
    agent.to_device(0)
    return agent


def _samples(g, name, itr):
    from rlpyt_b200.agents.pg.base import AgentInfoRnn
    from rlpyt_b200.distributions.categorical import DistInfo
    from rlpyt_b200.models.pg.atari_lstm_model import RnnState
    from rlpyt_b200.samplers.collections import AgentSamplesBsv, EnvSamples, Samples
    t = lambda k: torch.from_numpy(g[f"{name}/itr{itr}/{k}"]).cuda()
    all_action = torch.cat([torch.zeros(1, B, dtype=torch.int64, device="cuda"), t("action")])
    all_reward = torch.cat([torch.zeros(1, B, device="cuda"), t("reward")])
    return Samples(
        agent=AgentSamplesBsv(action=all_action[1:], prev_action=all_action[:-1],
                              agent_info=AgentInfoRnn(dist_info=DistInfo(prob=t("old_prob")), value=t("value"),
                                                      prev_rnn_state=RnnState(h=t("h0"), c=t("c0"))),
                              bootstrap_value=t("bv")),
        env=EnvSamples(observation=t("obs"), reward=all_reward[1:], prev_reward=all_reward[:-1], done=t("done"), env_info=None))


def test_recurrent_ppo_two_iterations_vs_reference(golden):
    """Same weights, samples, recorded initial rnn states and numpy shuffle stream as the reference's CPU run: OptInfo
    rows of the first update within 1e-5, later ones within 2e-4 (fp32 summation order through Adam; the LSTM runs on
    cuDNN here and on ATen-CPU there)."""
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.samplers.collections import BatchSpec
    g = golden("ppo_lstm")
    agent = _agent(g, "ppo")
    assert agent.recurrent
    algo = PPO(gae_lambda=0.95, minibatches=2, epochs=2)
    algo.initialize(agent, 4, BatchSpec(T, B), mid_batch_reset=True)
    np.random.seed(78)
    for itr in range(2):
        agent.train_mode(itr)
        info = algo.optimize_agent(itr, _samples(g, "ppo", itr))
        for f in ("loss", "gradNorm", "entropy", "perplexity"):
            got, want = np.asarray(getattr(info, f)), g[f"ppo/itr{itr}/opt_{f}"]
            assert got.shape == want.shape == (4,)
            if itr == 0:
                np.testing.assert_allclose(got[0], want[0], rtol=1e-5, atol=1e-7, err_msg=f)
            np.testing.assert_allclose(got, want, rtol=2e-4, atol=1e-6, err_msg=f)


def test_recurrent_a2c_iteration_vs_reference(golden):
    from rlpyt_b200.algos.pg.a2c import A2C
    from rlpyt_b200.samplers.collections import BatchSpec
    g = golden("ppo_lstm")
    agent = _agent(g, "a2c")
    algo = A2C(gae_lambda=0.95)
    algo.initialize(agent, 4, BatchSpec(T, B), mid_batch_reset=True)
    agent.train_mode(0)
    info = algo.optimize_agent(0, _samples(g, "a2c", 0))
    for f in ("loss", "gradNorm", "entropy", "perplexity"):
        np.testing.assert_allclose(getattr(info, f), g[f"a2c/itr0/opt_{f}"][0], rtol=2e-5, atol=1e-7, err_msg=f)


@pytest.mark.parametrize("kind", ["gpu", "alternating", "serial"])
def test_samplers_carry_the_recurrent_state(kind):
    """The samplers stay agnostic of the rnn state (agents/base.py:252-306): the recorded ``prev_rnn_state[t]`` is the
    state the agent started step t from - zeros at the first step and (standard samplers) in the column of an env
    right after its episode ended - and feeding the recorded inputs and states back through the model reproduces the
    recorded policy outputs; a sampler -> recurrent PPO iteration runs on the device."""
    from rlpyt_b200.agents.pg.atari import AlternatingAtariLstmAgent, AtariLstmAgent
    from rlpyt_b200.algos.pg.ppo import PPO
    from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
    from rlpyt_b200.samplers.parallel.gpu.alternating_sampler import AlternatingSampler
    from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
    from rlpyt_b200.samplers.serial.sampler import SerialSampler
    Tn, Bn = 10, 8
    cls = dict(gpu=GpuSampler, alternating=AlternatingSampler, serial=SerialSampler)[kind]
    sampler = cls(EnvCls=SyntheticAtariEnv, env_kwargs=dict(image_shape=IMAGE, n_actions=A, p_done=0.15, p_reward=0.3),
                  batch_T=Tn, batch_B=Bn, max_decorrelation_steps=0)
    Agent = AlternatingAtariLstmAgent if kind == "alternating" else AtariLstmAgent
    agent = Agent(model_kwargs=dict(fc_sizes=128, lstm_size=H))
    sampler.initialize(agent, affinity=dict(cuda_idx=0, workers_cpus=[None, None], set_affinity=False), seed=5, bootstrap_value=True)
    agent.to_device(0)
    algo = PPO(gae_lambda=0.95, minibatches=2, epochs=1)
    algo.initialize(agent, 10, sampler.batch_spec, mid_batch_reset=sampler.mid_batch_reset)
    try:
        for itr in range(3):
            agent.sample_mode(itr)
            samples, _ = sampler.obtain_samples(itr)
            st = samples.agent.agent_info.prev_rnn_state
            assert st.h.shape == (Tn, Bn, 1, H) and st.h.is_cuda
            if itr == 0:
                assert float(st.h[0].abs().max()) == 0.0 and float(st.c[0].abs().max()) == 0.0
            assert float(st.h[1:].abs().max()) > 0.0
            done = samples.env.done
            if kind != "alternating":          # the alternating mixin keeps its state across episode ends (as the reference's)
                for t in range(Tn - 1):
                    for b in torch.nonzero(done[t]).flatten().tolist():
                        assert float(st.h[t + 1, b].abs().max()) == 0.0 and float(st.c[t + 1, b].abs().max()) == 0.0
            # one-step consistency: model(obs[t], onehot(prev_action seen), prev_reward seen, state[t]) == recorded prob[t]
            t = 3
            pa = samples.agent.prev_action[t].clone()
            pr = samples.env.prev_reward[t].clone()
            if kind != "serial" and t > 0:     # the GPU action server zeroes the agent's inputs after an episode end
                pa[done[t - 1]] = 0
                pr[done[t - 1]] = 0
            with torch.no_grad():
                init = tuple(x.transpose(0, 1).contiguous() for x in (st.h[t], st.c[t]))
                pi, v, _ = agent.model(samples.env.observation[t], agent.distribution.to_onehot(pa), pr, init)
            np.testing.assert_allclose(pi.cpu().numpy(), samples.agent.agent_info.dist_info.prob[t].cpu().numpy(), rtol=1e-5, atol=1e-6)
            agent.train_mode(itr)
            info = algo.optimize_agent(itr, samples)
            assert len(info.loss) == 2 and all(np.isfinite(info.loss)) and all(np.isfinite(info.gradNorm))
    finally:
        sampler.shutdown()
