"""Asynchronous runner on the device (reference rlpyt/runners/async_rl.py): the real
asynchronous samplers (serial / parallel / alternating) + AtariDqnAgent + DQN with the lock-and-fence guarded HBM replay,
driven by ``AsyncRl.train()``.  The data-integrity check records every batch the sampler publishes and compares the
replay ring with their concatenation - a torn or reordered hand-off (missing stream ordering between the sampler's,
the copier's and the optimizer's CUDA streams) would show up there."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

IMG, A = (4, 84, 84), 6


def _build(kind, T=4, B=8, n_itr=24, **algo_kw):
    from rlpyt_b200.agents.dqn.atari.atari_dqn_agent import AtariDqnAgent
    from rlpyt_b200.algos.dqn.dqn import DQN
    from rlpyt_b200.envs.synthetic import SyntheticAtariEnv
    from rlpyt_b200.runners.async_rl import AsyncRl, AsyncRlEval
    from rlpyt_b200.samplers.async_ import AsyncAlternatingSampler, AsyncGpuSampler, AsyncSerialSampler
    from rlpyt_b200.utils.logging import TabularLogger
    cls = dict(serial=AsyncSerialSampler, gpu=AsyncGpuSampler, alternating=AsyncAlternatingSampler, eval=AsyncSerialSampler)[kind]
    kw = dict(eval_n_envs=2, eval_max_steps=40, eval_max_trajectories=4) if kind == "eval" else {}
    sampler = cls(EnvCls=SyntheticAtariEnv, env_kwargs=dict(image_shape=IMG, n_actions=A, p_done=0.05, p_reward=0.3),
                  batch_T=T, batch_B=B, max_decorrelation_steps=0, **kw)
    args = dict(batch_size=32, min_steps_learn=2 * T * B, replay_size=64 * T * B, replay_ratio=8, n_step_return=3,
                double_dqn=True, prioritized_replay=True, target_update_interval=4, updates_per_sync=2)
    args.update(algo_kw)
    algo = DQN(**args)
    agent = AtariDqnAgent()
    logger = TabularLogger(quiet=True)
    runner_cls = AsyncRlEval if kind == "eval" else AsyncRl
    runner = runner_cls(algo=algo, agent=agent, sampler=sampler, n_steps=n_itr * T * B,
                        affinity=dict(cuda_idx=0, workers_cpus=[None, None], set_affinity=False), seed=5,
                        log_interval_steps=(n_itr // 3) * T * B, logger=logger)
    runner.throttle_wait = 0.002
    return runner, sampler, algo, agent, logger


@pytest.mark.parametrize("kind", ["serial", "gpu", "alternating"])
def test_async_rl_trains_dqn_and_replay_holds_exactly_what_the_sampler_published(kind):
    runner, sampler, algo, agent, logger = _build(kind)
    published = []
    orig = sampler.obtain_samples

    def recording(itr, db_idx):
        out = orig(itr, db_idx)
        torch.cuda.current_stream().synchronize()
        db = sampler.double_buffer[db_idx]
        published.append(dict(action=db.agent.action.cpu().numpy().copy(), reward=db.env.reward.cpu().numpy().copy(),
                              done=db.env.done.cpu().numpy().copy(),
                              frame=db.env.observation[:, :, -1].cpu().numpy().copy()))
        return out
    sampler.obtain_samples = recording
    n_opt = runner.train()
    n_itr, T, B = runner.n_itr, sampler.batch_spec.T, sampler.batch_spec.B
    assert len(published) == n_itr and n_opt > 0 and algo.update_counter == n_opt * algo.updates_per_optimize
    rb = algo.replay_buffer
    assert rb.async_ and rb.t == n_itr * T and not rb._buffer_full           # ring not wrapped: contents comparable
    cat = lambda k: np.concatenate([p[k] for p in published])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(rb.samples.action[:rb.t].cpu().numpy(), cat("action"))
    np.testing.assert_array_equal(rb.samples.reward[:rb.t].cpu().numpy(), cat("reward"))
    np.testing.assert_array_equal(rb.samples.done[:rb.t].cpu().numpy(), cat("done"))
    np.testing.assert_array_equal(rb.samples_new_frames[:rb.t].cpu().numpy(), cat("frame"))
    # sum-tree still consistent after interleaved advances (copier stream) and priority updates (optimizer stream)
    tree = rb.priority_tree
    nodes = tree.tree.cpu().numpy()
    np.testing.assert_allclose(nodes[0], nodes[tree.low_idx:tree.high_idx].sum(), rtol=1e-12)
    # the optimizer respected the replay-ratio bound and its parameters reached the sampler's copy
    cum_replay_ratio = algo.update_counter * algo.batch_size / ((n_itr - 1) * T * B)
    assert cum_replay_ratio <= algo.replay_ratio * 1.05 + algo.batch_size * algo.updates_per_optimize / (T * B)
    twin = sampler.agent
    assert twin is not agent and twin.model is not agent.model
    assert agent._async["send_count"] == n_opt               # sent after every optimize_agent (a fast sampler may finish before the first)
    twin.recv_shared_memory()
    torch.cuda.synchronize()
    for a, b in zip(agent.model.state_dict().values(), twin.model.state_dict().values()):
        assert torch.equal(a, b)                                              # last send == current parameters (sent after every optimize_agent)
    t = logger.tables[-1]
    assert t["Diagnostics/CumUpdates"] == algo.update_counter
    assert any(np.isfinite(tb.get("lossAverage", np.nan)) for tb in logger.tables)   # (a table whose interval saw no update logs NaN, like the reference)
    assert t["Diagnostics/CumSteps"] == (n_itr - 1) * T * B


def test_async_rl_eval_runs_offline_evaluation_in_the_sampler_thread():
    runner, sampler, algo, agent, logger = _build("eval", n_itr=12)
    runner.train()
    assert logger.tables[0]["Diagnostics/TrajsInEval"] >= 1 and logger.tables[-1]["Diagnostics/CumEvalTime"] > 0
    assert algo.update_counter > 0


def test_async_replay_stream_fence_orders_append_against_sampling():
    """Two host threads, two streams, no host synchronisation in between: every sampled row must be a row that some
    append wrote completely (value pattern: every field of row t carries the same batch number)."""
    import threading
    from rlpyt_b200.algos.dqn.dqn import SamplesToBuffer
    from rlpyt_b200.replays.non_sequence.frame import AsyncUniformReplayFrameBuffer
    T, B, n_batches = 8, 4, 40
    ex = SamplesToBuffer(observation=np.zeros(IMG, np.uint8), action=np.int64(0), reward=np.float32(0), done=np.bool_(False))
    rb = AsyncUniformReplayFrameBuffer(example=ex, size=T * B * (n_batches + 2), B=B, n_step_return=1, device=torch.device("cuda", 0))
    errors = []

    def writer():
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                for k in range(1, n_batches + 1):
                    obs = torch.full((T, B) + IMG, k % 251, dtype=torch.uint8, device="cuda")
                    rb.append_samples(SamplesToBuffer(observation=obs, action=torch.full((T, B), k, device="cuda"),
                                                      reward=torch.full((T, B), float(k), device="cuda"),
                                                      done=torch.zeros(T, B, dtype=torch.bool, device="cuda")))
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    th = threading.Thread(target=writer)
    th.start()
    checked = 0
    with torch.cuda.stream(torch.cuda.Stream()):
        while th.is_alive() or checked < 5:
            if rb.t < 2 * T:
                continue
            batch = rb.sample_batch(16)
            a = batch.action.cpu().numpy()
            r = batch.return_.cpu().numpy()
            f = batch.agent_inputs.observation[:, -1, 0, 0].cpu().numpy()
            assert np.all(a >= 1) and np.array_equal(r, a.astype(np.float32)) and np.array_equal(f, (a % 251).astype(np.uint8))
            checked += 1
    th.join()
    assert not errors and checked >= 5
