"""Shared driver for the replay parity tests: replays the recorded append / sample / update stream
of tests/golden/replay.npz (made by the reference's PrioritizedReplayFrameBuffer /
UniformReplayFrameBuffer) through any buffer implementation with the oracle's interface."""
import numpy as np

CASES = ["kat", "small_pri", "small_pri_unique", "small_uni", "n1_f1", "mid_pri", "bigT_append"]
FIELDS = ["observation", "prev_action", "prev_reward", "action", "return_", "done", "done_n",
          "target_observation", "target_prev_action", "target_prev_reward"]


def replay_stream(seed, n_batches, T, B, obs_shape, A, p_done):
    """Must stay identical to tests/golden/make_golden.py:replay_stream."""
    rng = np.random.default_rng(seed)
    nf = obs_shape[0]
    hist = rng.integers(0, 256, size=(nf - 1, B) + tuple(obs_shape[1:]), dtype=np.uint8)
    for _ in range(n_batches):
        new = rng.integers(0, 256, size=(T, B) + tuple(obs_shape[1:]), dtype=np.uint8)
        full = np.concatenate([hist, new], 0)
        obs = np.stack([full[c:c + T] for c in range(nf)], axis=2)
        hist = full[-(nf - 1):] if nf > 1 else hist
        yield dict(observation=obs, action=rng.integers(0, A, size=(T, B)).astype(np.int64),
                   reward=rng.standard_normal((T, B)).astype(np.float32), done=rng.random((T, B)) < p_done)


def case_config(g, name):
    seed, size, B, n_step, batch_T, n_batches, batch_B, prioritized, unique = (int(x) for x in g[f"{name}/cfg"])
    return dict(seed=seed, size=size, B=B, n_step=n_step, batch_T=batch_T, n_batches=n_batches, batch_B=batch_B,
                prioritized=bool(prioritized), unique=bool(unique), obs_shape=tuple(int(x) for x in g[f"{name}/obs_shape"]),
                discount=float(g[f"{name}/discount"][0]))


def drive(g, name, make_buffer, to_np=np.asarray, check_root=None):
    """``make_buffer(cfg)`` -> object with append_samples / sample_batch(batch_B, random_values=) /
    update_batch_priorities / t.  Asserts bit-equality with the golden record at every step."""
    c = case_config(g, name)
    buf = make_buffer(c)
    np.random.seed(c["seed"])
    rng = np.random.default_rng(c["seed"] + 1000)
    n_checked = 0
    for i, s in enumerate(replay_stream(c["seed"], c["n_batches"], c["batch_T"], c["B"], c["obs_shape"], 4, 0.1)):
        buf.append_samples(s)
        assert int(buf.t) == int(g[f"{name}/b{i}/t"][0])
        if c["prioritized"] and check_root is not None:
            assert check_root(buf) == g[f"{name}/b{i}/root"][0], (name, i)
        if f"{name}/b{i}/action" not in g.files:  # nothing sampled at this point of the record
            continue
        u = rng.random(c["batch_B"])
        if c["prioritized"] and not c["unique"]:
            assert np.array_equal(u, g[f"{name}/b{i}/uniforms"])
            batch = buf.sample_batch(c["batch_B"], random_values=u)
        else:
            batch = buf.sample_batch(c["batch_B"])
        for k in FIELDS:
            got, want = to_np(batch[k]), g[f"{name}/b{i}/{k}"]
            assert got.dtype == want.dtype and got.shape == want.shape, (name, i, k, got.dtype, want.dtype)
            assert np.array_equal(got, want), (name, i, k)
        if c["prioritized"]:
            tree_idxs = to_np(batch["T_idxs"]) * c["B"] + to_np(batch["B_idxs"])
            low = g[f"{name}/b{i}/tree_idxs"] - (to_np(batch["T_idxs"]) * c["B"] + to_np(batch["B_idxs"]))
            assert np.all(low == low[0])
            assert np.array_equal(to_np(batch["is_weights"]), g[f"{name}/b{i}/is_weights"])
            new_pri = g[f"{name}/b{i}/new_pri"]
            assert np.array_equal(np.abs(rng.standard_normal(c["batch_B"])).astype(np.float32) + 0.01, new_pri)
            buf.update_batch_priorities(new_pri)
            if check_root is not None:
                assert check_root(buf) == g[f"{name}/b{i}/root_after"][0], (name, i)
        else:
            assert np.array_equal(to_np(batch["T_idxs"]), g[f"{name}/b{i}/T_idxs"])
            assert np.array_equal(to_np(batch["B_idxs"]), g[f"{name}/b{i}/B_idxs"])
        n_checked += 1
    assert n_checked >= 2
    return buf


SEQ_CASES = ["seq_uni_norn", "seq_uni_rsi1", "seq_uni_rsi4", "seq_pri_rsi1", "seq_pri_rsi4_input", "seq_pri_norn", "seq_mid"]


def seq_replay_case(g, name, make, append, sample, update):
    """Drives one recorded stream through ``make/append/sample/update`` and compares every recorded output."""
    (seed, size, B, n_step, sampler_T, n_batches, batch_B, batch_T, rsi, prioritized, input_pri, pri_shift) = \
        (int(v) for v in g[f"{name}/cfg"])
    obs_shape = tuple(int(v) for v in g[f"{name}/obs_shape"])
    buf = make(obs_shape=obs_shape, size=size, B=B, rsi=rsi, batch_T=batch_T, discount=float(g[f"{name}/discount"][0]),
               n_step=n_step, prioritized=bool(prioritized), input_pri=bool(input_pri), pri_shift=pri_shift)
    assert buf.T == int(g[f"{name}/T"][0])
    np.random.seed(seed)
    n_sampled = 0
    for i, s in enumerate(replay_stream(seed, n_batches, sampler_T, B, obs_shape, 4, 0.08)):
        if rsi > 0:
            s["prev_rnn_state"] = dict(h=g[f"{name}/b{i}/rnn_h"], c=g[f"{name}/b{i}/rnn_c"])
        append(buf, s, g[f"{name}/b{i}/input_pri"] if input_pri else None)
        assert buf.t == int(g[f"{name}/b{i}/t"][0])
        if f"{name}/b{i}/all_action" not in g:
            continue
        n_sampled += 1
        batch = sample(buf, batch_B, g[f"{name}/b{i}/uniforms"] if prioritized else None)
        if prioritized:
            assert np.array_equal(batch["is_weights"], g[f"{name}/b{i}/is_weights"]), (name, i)
        else:
            assert np.array_equal(batch["T_idxs"], g[f"{name}/b{i}/T_idxs"])
            assert np.array_equal(batch["B_idxs"], g[f"{name}/b{i}/B_idxs"])
        for k in ("all_observation", "all_action", "all_reward", "return_", "done", "done_n"):
            assert np.array_equal(batch[k], g[f"{name}/b{i}/{k}"]), (name, i, k)
        if rsi > 0:
            assert np.array_equal(batch["init_rnn_state"]["h"], g[f"{name}/b{i}/init_h"])
            assert np.array_equal(batch["init_rnn_state"]["c"], g[f"{name}/b{i}/init_c"])
        if prioritized:
            update(buf, g[f"{name}/b{i}/new_pri"])
    assert n_sampled >= 5
    return buf
