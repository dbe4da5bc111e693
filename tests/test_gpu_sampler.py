"""Collector / sampler tests on the GPU box: the parallel GpuSampler (forked env workers + device-
resident batch) and the SerialSampler fill the [T,B] buffers with exactly what the envs and the
agent produced (checked against a host-side replay of the same seeded envs), keep the
prev_action/prev_reward aliasing of rlpyt/samplers/buffer.py:29-45, and drive a full
sampler -> PPO iteration."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(sampler_cls, T=6, B=8, n_workers=2, image=(4, 36, 36), collector=None, decor=0, p_done=0.1, p_reward=0.3):
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
    kw = dict(EnvCls=SyntheticAtariEnv, env_kwargs=dict(image_shape=image, n_actions=5, p_done=p_done, p_reward=p_reward),
              batch_T=T, batch_B=B, max_decorrelation_steps=decor)
    if collector is not None:
        kw["CollectorCls"] = collector
    sampler = sampler_cls(**kw)
    agent = AtariFfAgent()
    aff = dict(cuda_idx=0, workers_cpus=[None] * n_workers, set_affinity=False)
    sampler.initialize(agent, affinity=aff, seed=11, bootstrap_value=True)
    agent.to_device(0)
    return sampler, agent


@pytest.mark.parametrize("kind,image,B,p_done,p_reward", [
    ("gpu", (4, 36, 36), 8, 0.1, 0.3), ("serial", (4, 36, 36), 8, 0.1, 0.3), ("alternating", (4, 36, 36), 8, 0.1, 0.3),
    # full-size frames, every step rewarded, half of the steps end an episode: the reward / done of the LAST step of a
    # batch must be recorded before the master zeroes the step buffer for the next one (round-1 race)
    ("gpu", (4, 84, 84), 16, 0.5, 1.0), ("alternating", (4, 84, 84), 16, 0.5, 1.0)])
def test_sampler_batch_matches_env_replay(kind, image, B, p_done, p_reward):
    """Every sampler fills the [T,B] buffers with exactly what the seeded envs produce under the recorded actions -
    eager first batch and CUDA-graph replays alike (4 iterations); the alternating sampler (two worker groups,
    two half-batch step engines) therefore yields the same batch layout as the standard one."""
    from rlpyt_b200.samplers.parallel.gpu.alternating_sampler import AlternatingSampler
    from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
    from rlpyt_b200.samplers.serial.sampler import SerialSampler
    from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
    T = 6
    cls = dict(gpu=GpuSampler, serial=SerialSampler, alternating=AlternatingSampler)[kind]
    sampler, agent = _mk(cls, T, B, image=image, p_done=p_done, p_reward=p_reward)
    try:
        # host-side replay: same seeds (worker w gets seed+w, env i of a worker +i; serial: seed+i)
        if kind != "serial":
            seeds = [11 + w + i for w in range(2) for i in range(B // 2)]
        else:
            seeds = [11 + i for i in range(B)]
        envs = [SyntheticAtariEnv(image_shape=image, n_actions=5, p_done=p_done, p_reward=p_reward) for _ in range(B)]
        obs = []
        for e, s in zip(envs, seeds):
            e.seed(s)
            obs.append(e.reset())
        n_done_total = n_infos_total = 0
        for itr in range(4):
            samples, traj_infos = sampler.obtain_samples(itr)
            s_obs = samples.env.observation.cpu().numpy()
            s_act = samples.agent.action.cpu().numpy()
            s_rew = samples.env.reward.cpu().numpy()
            s_done = samples.env.done.cpu().numpy()
            n_done = 0
            for t in range(T):
                for b, e in enumerate(envs):
                    # serial: env 0 also served build_samples_buffer's example step (as in the
                    # reference, serial/sampler.py:58-60), so its RNG stream is offset - not replayed.
                    if kind == "serial" and b == 0:
                        continue
                    assert np.array_equal(s_obs[t, b], obs[b]), (itr, t, b)
                    o, r, d, info = e.step(s_act[t, b])
                    assert r == s_rew[t, b] and d == s_done[t, b]
                    if d:
                        o = e.reset()
                        n_done += 1
                    obs[b] = o
            n_done_total += n_done
            n_infos_total += len(traj_infos)
            # completed TrajInfos travel through a multiprocessing queue (feeder thread): the ones of the last steps
            # may surface at the next obtain_samples, never more than were completed (parallel/base.py:104-113)
            assert kind == "serial" or n_infos_total <= n_done_total
            # aliasing: prev_action[t+1] is action[t]; prev_reward likewise (buffer.py:29-45)
            assert torch.equal(samples.agent.prev_action[1:], samples.agent.action[:-1])
            assert torch.equal(samples.env.prev_reward[1:], samples.env.reward[:-1])
            prob = samples.agent.agent_info.dist_info.prob
            assert prob.is_cuda and samples.env.observation.is_cuda
            np.testing.assert_allclose(prob.sum(-1).cpu().numpy(), 1.0, rtol=1e-5)
            # recorded prob/value are what the network gives on the recorded observations
            with torch.no_grad():
                pi, v = agent.model(samples.env.observation, None, None)
            np.testing.assert_allclose(prob.cpu().numpy(), pi.cpu().numpy(), rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(samples.agent.agent_info.value.cpu().numpy(), v.cpu().numpy(), rtol=1e-5, atol=1e-6)
            assert samples.agent.bootstrap_value.shape == (1, B)
    finally:
        sampler.shutdown()


def test_wait_reset_collector_blanks_after_done():
    from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
    from rlpyt_b200.samplers.collectors import GpuWaitResetCollector
    from rlpyt_b200.algos.utils import valid_from_done
    T, B = 12, 8
    sampler, agent = _mk(GpuSampler, T, B, collector=GpuWaitResetCollector, p_done=0.2)
    try:
        assert sampler.mid_batch_reset is False
        for itr in range(2):
            samples, _ = sampler.obtain_samples(itr)
            done = samples.env.done.cpu().numpy()
            valid = valid_from_done(samples.env.done).cpu().numpy()
            for b in range(B):
                if done[:, b].any():
                    first = int(np.argmax(done[:, b]))
                    assert done[first:, b].all()                      # done stays True to the end
                    dead = slice(first + 1, T)
                    assert (samples.agent.action[dead, b] == 0).all()
                    assert (samples.env.reward[dead, b] == 0).all()
                    assert (samples.env.observation[dead, b] == 0).all()
                    assert (samples.agent.agent_info.value[dead, b] == 0).all()
                    assert (valid[: first + 1, b] == 1).all() and (valid[first + 1:, b] == 0).all()
    finally:
        sampler.shutdown()


def test_sampler_to_ppo_iteration_runs_on_device():
    from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
    from rlpyt_b200.algos.pg.ppo import PPO
    sampler, agent = _mk(GpuSampler, T=16, B=8, decor=5)
    try:
        algo = PPO(gae_lambda=0.98, minibatches=2, epochs=2)
        algo.initialize(agent, 10, sampler.batch_spec, mid_batch_reset=sampler.mid_batch_reset)
        for itr in range(3):
            samples, _ = sampler.obtain_samples(itr)
            agent.train_mode(itr)
            info = algo.optimize_agent(itr, samples)
            assert len(info.loss) == 4 and all(np.isfinite(info.loss)) and all(np.isfinite(info.gradNorm))
            agent.sample_mode(itr)
    finally:
        sampler.shutdown()


def test_config1_serial_a2c_cartpole():
    """BASELINE.json config 1: SerialSampler + A2C on CartPole, T=5, B=8 (plumbing)."""
    from rlpyt_b200.samplers.serial.sampler import SerialSampler
    from rlpyt_b200.algos.pg.a2c import A2C
    from rlpyt_b200.agents.pg.categorical import CategoricalPgAgent
    from rlpyt_b200.envs.cartpole import CartPoleEnv
    from rlpyt_b200.models.mlp import MlpModel

    class CartPoleModel(torch.nn.Module):
        def __init__(self, obs_dim, n_actions):
            super().__init__()
            self.body = MlpModel(obs_dim, [64, 64])
            self.pi = torch.nn.Linear(64, n_actions)
            self.v = torch.nn.Linear(64, 1)

        def forward(self, observation, prev_action, prev_reward):
            lead = observation.dim() - 1
            x = self.body(observation.reshape(-1, observation.shape[-1]).float())
            pi = torch.softmax(self.pi(x), -1).reshape(observation.shape[:lead] + (-1,))
            return pi, self.v(x).reshape(observation.shape[:lead])

    class Agent(CategoricalPgAgent):
        def make_env_to_model_kwargs(self, env_spaces):
            return dict(obs_dim=env_spaces.observation.shape[0], n_actions=env_spaces.action.n)

    sampler = SerialSampler(EnvCls=CartPoleEnv, env_kwargs=dict(), batch_T=5, batch_B=8, max_decorrelation_steps=0)
    agent = Agent(ModelCls=CartPoleModel)
    sampler.initialize(agent, affinity=dict(cuda_idx=0), seed=0, bootstrap_value=True)
    agent.to_device(0)
    algo = A2C()
    algo.initialize(agent, 20, sampler.batch_spec, mid_batch_reset=True)
    for itr in range(20):
        samples, traj_infos = sampler.obtain_samples(itr)
        agent.train_mode(itr)
        info = algo.optimize_agent(itr, samples)
        assert np.isfinite(info.loss) and np.isfinite(info.gradNorm)
    assert algo.update_counter == 20


@pytest.mark.parametrize("kind", ["gpu", "alternating", "gpu_wait_reset", "serial", "gpu+chunked", "alternating+chunked"])
def test_samplers_match_reference_sampler_golden(kind, golden, monkeypatch):
    """Field-by-field parity with the REFERENCE's samplers (tests/golden/collector.npz: rlpyt's GpuSampler,
    GpuSampler + GpuWaitResetCollector and SerialSampler stepping the same seeded synthetic envs under the
    deterministic policy of tests/deterministic_agent.py, three consecutive batches): observations, actions, rewards,
    done flags, the prev_action / prev_reward views, recorded agent_info, bootstrap values and env_info - including
    what the agent saw as previous action / reward after an episode end (zeroed by the GPU action server, kept by
    the CPU collector: SURVEY.md 9.5) since the policy depends on both.  The alternating sampler must produce the
    standard GPU sampler's batch."""
    import os
    import sys
    if kind.endswith("+chunked"):        # per-worker observation uploads (rl_upload_async) + polling master
        kind = kind[:-len("+chunked")]
        monkeypatch.setenv("RLPYT_B200_SAMPLER_CHUNKED", "1")
        monkeypatch.setenv("RLPYT_B200_SAMPLER_POLL", "spin")
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from deterministic_agent import make_agent_class
    from rlpyt_b200.agents.base import AgentStep
    from rlpyt_b200.agents.pg.base import AgentInfo
    from rlpyt_b200.distributions.categorical import DistInfo
    from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
    from rlpyt_b200.samplers.collectors import GpuWaitResetCollector
    from rlpyt_b200.samplers.parallel.gpu.alternating_sampler import AlternatingSampler
    from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
    from rlpyt_b200.samplers.serial.sampler import SerialSampler
    g = golden("collector")
    T, B, A = int(g["T"][0]), int(g["B"][0]), int(g["A"][0])
    image, seed = tuple(int(x) for x in g["image"]), int(g["seed"][0])
    env_kwargs = dict(image_shape=image, n_actions=A, p_done=float(g["p_done"][0]), p_reward=float(g["p_reward"][0]))
    cls = dict(gpu=GpuSampler, alternating=AlternatingSampler, gpu_wait_reset=GpuSampler, serial=SerialSampler)[kind]
    extra = dict(CollectorCls=GpuWaitResetCollector) if kind == "gpu_wait_reset" else {}
    sampler = cls(EnvCls=SyntheticAtariEnv, env_kwargs=env_kwargs, batch_T=T, batch_B=B, max_decorrelation_steps=0, **extra)
    agent = make_agent_class(AgentStep, AgentInfo, DistInfo)()
    sampler.initialize(agent, affinity=dict(cuda_idx=0, workers_cpus=[None, None], set_affinity=False), seed=seed,
                       bootstrap_value=True, traj_info_kwargs=dict(discount=0.99))
    agent.to_device(0)
    ref = "gpu" if kind == "alternating" else kind
    cols = slice(1, B) if kind == "serial" else slice(0, B)   # serial: env 0 also served the example step (see test_oracle_collector)
    try:
        for itr in range(3):
            agent.sample_mode(itr)
            samples, traj_infos = sampler.obtain_samples(itr)
            pre = f"{ref}/itr{itr}/"
            got = {
                "observation": samples.env.observation, "reward": samples.env.reward, "prev_reward": samples.env.prev_reward,
                "done": samples.env.done, "action": samples.agent.action, "prev_action": samples.agent.prev_action,
                "prob": samples.agent.agent_info.dist_info.prob, "value": samples.agent.agent_info.value,
                "bootstrap_value": samples.agent.bootstrap_value,
                "traj_done": samples.env.env_info.traj_done, "game_score": samples.env.env_info.game_score,
            }
            for k, v in got.items():
                a = v.cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
                want = g[pre + k]
                assert a.shape == want.shape, (k, a.shape, want.shape)
                assert np.array_equal(a[:, cols], want[:, cols]), (kind, itr, k)
            if kind != "serial":
                assert len(traj_infos) == int(g[pre + "n_traj"][0])
                assert sorted(int(t["Length"]) for t in traj_infos) == list(g[pre + "traj_lengths"])
    finally:
        sampler.shutdown()



def test_first_layer_streams_frames_from_pinned_host_memory():
    """csrc/conv1_i8.cuh ``copy_out`` (rl_conv1_u8_forward_i8_stream): frames read by the kernel's bulk copies straight
    out of page-locked host memory give the same activations, bit for bit, as the same frames resident in HBM, and land
    in the HBM copy unchanged."""
    from rlpyt_b200.models import conv1_op
    from rlpyt_b200.utils.gather import HostMappedFrames
    g = torch.Generator().manual_seed(5)
    for B in (1, 37, 128, 300):
        host = torch.randint(0, 256, (B, 4, 84, 84), dtype=torch.uint8, generator=g).pin_memory()
        w = (torch.randn(16, 4, 8, 8, generator=g) / 16).cuda()
        b = torch.randn(16, generator=g).cuda()
        copy = torch.zeros((B, 4, 84, 84), dtype=torch.uint8, device="cuda")
        got = conv1_op.conv1_u8_relu_stream(w, b, HostMappedFrames(host.data_ptr(), copy))
        want = conv1_op.conv1_u8_relu(w, b, host.cuda(), None)
        torch.cuda.synchronize()
        assert torch.equal(got, want), B
        assert torch.equal(copy.cpu(), host), B


@pytest.mark.parametrize("kind", ["gpu", "alternating", "serial"])
def test_zero_copy_sampler_steps_match_the_uploading_sampler(kind, monkeypatch):
    """RLPYT_B200_SAMPLER_ZEROCOPY=1 (agent.step's first layer reads the step buffer over PCIe and records
    observation[t] itself, no H2D in front of it) leaves every field of three consecutive batches - the eager first
    batch and the CUDA-graph replays - identical to the sampler that uploads first.  Same seeds, same Philox stream."""
    from rlpyt_b200.agents.pg.atari import AtariFfAgent
    from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
    from rlpyt_b200.samplers.parallel.gpu.alternating_sampler import AlternatingSampler
    from rlpyt_b200.samplers.parallel.gpu.sampler import GpuSampler
    from rlpyt_b200.samplers.serial.sampler import SerialSampler
    cls = dict(gpu=GpuSampler, alternating=AlternatingSampler, serial=SerialSampler)[kind]
    T, B = 6, 8

    def run(zero_copy):
        monkeypatch.setenv("RLPYT_B200_SAMPLER_ZEROCOPY", "1" if zero_copy else "0")
        torch.manual_seed(0)
        sampler = cls(EnvCls=SyntheticAtariEnv, env_kwargs=dict(image_shape=(4, 84, 84), n_actions=6, p_done=0.1, p_reward=0.3),
                      batch_T=T, batch_B=B, max_decorrelation_steps=0)
        agent = AtariFfAgent()
        sampler.initialize(agent, affinity=dict(cuda_idx=0, workers_cpus=[None] * 4, set_affinity=False), seed=4, bootstrap_value=True)
        agent.to_device(0)
        ros = getattr(sampler, "rollouts", None) or [sampler.rollout]
        assert all(ro.zero_copy == zero_copy for ro in ros)
        out = []
        try:
            for itr in range(3):
                samples, _ = sampler.obtain_samples(itr)
                torch.cuda.synchronize()
                out.append(dict(obs=samples.env.observation.cpu().clone(), action=samples.agent.action.cpu().clone(),
                                prob=samples.agent.agent_info.dist_info.prob.cpu().clone(), value=samples.agent.agent_info.value.cpu().clone(),
                                reward=samples.env.reward.cpu().clone(), done=samples.env.done.cpu().clone(),
                                bv=samples.agent.bootstrap_value.cpu().clone()))
        finally:
            sampler.shutdown()
        return out

    a, b = run(False), run(True)
    for itr, (x, y) in enumerate(zip(a, b)):
        for k in x:
            assert torch.equal(x[k], y[k]), (itr, k)
