"""GPU parity of the sequence (R2D1) replay classes: the streams recorded from the reference's
UniformSequenceReplayFrameBuffer / PrioritizedSequenceReplayFrameBuffer (tests/golden/seq_replay.npz) must come back
BIT-EXACTLY through rlpyt_b200.replays.sequence - indices, frame-stack sequences with blanking, action / reward
sequences (incl. the reference's placement of a start index of -1), n-step returns, stored RNN states, the whole
fp64 tree after input priorities and updates; importance weights (fp32 from an fp64 pow) to 1e-6."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from replay_cases import SEQ_CASES, seq_replay_case  # noqa: E402


def to_np(x):
    return x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def make(obs_shape, size, B, rsi, batch_T, discount, n_step, prioritized, input_pri, pri_shift):
    from rlpyt_b200.replays.sequence.frame import PrioritizedSequenceReplayFrameBuffer, UniformSequenceReplayFrameBuffer
    from rlpyt_b200.utils.collections import namedarraytuple
    fields = ["observation", "action", "reward", "done"] + (["prev_rnn_state"] if rsi > 0 else [])
    Example = namedarraytuple("SamplesToBufferSeq", fields)
    Rnn = namedarraytuple("RnnState", ["h", "c"])
    vals = dict(observation=np.zeros(obs_shape, np.uint8), action=np.int64(0), reward=np.float32(0), done=np.bool_(False))
    if rsi > 0:
        vals["prev_rnn_state"] = Rnn(h=np.zeros((1, 3), np.float32), c=np.zeros((1, 3), np.float32))
    kw = dict(example=Example(**vals), size=size, B=B, discount=discount, n_step_return=n_step, rnn_state_interval=rsi,
              batch_T=batch_T)
    if prioritized:
        buf = PrioritizedSequenceReplayFrameBuffer(alpha=0.6, beta=0.9, default_priority=1, input_priorities=input_pri,
                                                   input_priority_shift=pri_shift, pow_on_host=True, **kw)
    else:
        buf = UniformSequenceReplayFrameBuffer(**kw)
    buf._test = (Example, Rnn, namedarraytuple("PrioritiesSamplesToBuffer", ["priorities", "samples"]), prioritized)
    return buf


def append(buf, s, pri):
    Example, Rnn, Pri, _ = buf._test
    vals = {k: torch.from_numpy(v).cuda() for k, v in s.items() if k != "prev_rnn_state"}
    if "prev_rnn_state" in s:
        vals["prev_rnn_state"] = Rnn(**{k: torch.from_numpy(v).cuda() for k, v in s["prev_rnn_state"].items()})
    samples = Example(**vals)
    buf.append_samples(samples if pri is None else Pri(priorities=pri, samples=samples))


def sample(buf, n, uniforms):
    if buf._test[3]:
        b = buf.sample_batch(n, random_values=uniforms)
        out = dict(is_weights=to_np(b.is_weights))
    else:
        st = np.random.get_state()
        T_idxs, B_idxs = buf.sample_idxs(n, buf.batch_T)
        np.random.set_state(st)
        b = buf.sample_batch(n)
        out = dict(T_idxs=T_idxs, B_idxs=B_idxs)
    for k in ("all_observation", "all_action", "all_reward", "return_", "done", "done_n"):
        out[k] = to_np(getattr(b, k))
    if b.init_rnn_state is not None:
        out["init_rnn_state"] = dict(h=to_np(b.init_rnn_state.h), c=to_np(b.init_rnn_state.c))
    return out


@pytest.mark.parametrize("name", SEQ_CASES)
def test_sequence_replay_stream_bit_exact_on_gpu(golden, name, monkeypatch):
    import replay_cases
    g = golden("seq_replay")
    real_equal = np.array_equal

    def equal(a, b):  # fp32 importance weights: 1e-6 relative; everything else exact
        a, b = np.asarray(a), np.asarray(b)
        if a.dtype == np.float32 and a.ndim == 1 and a.shape == b.shape and a.size and (a.max() == 1.0 == b.max()):
            return np.allclose(a, b, rtol=1e-6, atol=0)
        return real_equal(a, b)
    monkeypatch.setattr(replay_cases.np, "array_equal", equal)
    buf = seq_replay_case(g, name, make, append, sample, lambda b, p: b.update_batch_priorities(torch.from_numpy(p).cuda()))
    monkeypatch.undo()
    if buf._test[3]:
        assert np.array_equal(buf.priority_tree.tree.cpu().numpy(), g[f"{name}/final_tree"])
    assert np.array_equal(buf.samples_return_.cpu().numpy(), g[f"{name}/final_return"])
    assert np.array_equal(buf.samples_done_n.cpu().numpy(), g[f"{name}/final_done_n"])
    if buf.rnn_state_interval > 1:
        assert np.array_equal(buf.samples_prev_rnn_state.h.cpu().numpy(), g[f"{name}/final_rnn_h"])


def test_sequence_extract_at_r2d1_scale_matches_gather():
    """batch_B=64 sequences of 40+80 (+5) steps of (4,84,84) frames: the kernel's output equals plain torch indexing of
    the frame ring (no done in the window => no blanking), and sequences that cross the ring's end wrap."""
    from rlpyt_b200.replays.sequence.frame import UniformSequenceReplayFrameBuffer
    from rlpyt_b200.utils.collections import namedarraytuple
    Example = namedarraytuple("SamplesToBufferBig", ["observation", "action", "reward", "done"])
    ex = Example(observation=np.zeros((4, 84, 84), np.uint8), action=np.int64(0), reward=np.float32(0), done=np.bool_(False))
    B, T = 16, 400
    buf = UniformSequenceReplayFrameBuffer(example=ex, size=T * B, B=B, discount=0.997, n_step_return=5, rnn_state_interval=0,
                                           batch_T=120)
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(T // 40 + 1):
        obs = torch.randint(0, 256, (40, B, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
        buf.append_samples(Example(observation=obs, action=torch.randint(0, 6, (40, B), device="cuda", generator=g),
                                   reward=torch.randn(40, B, device="cuda", generator=g),
                                   done=torch.zeros(40, B, dtype=torch.bool, device="cuda")))
    T_idxs = np.array([0, 41, 250, 279, 399, 390, 300, 120] * 8)
    B_idxs = np.arange(64) % B
    batch = buf.extract_batch(T_idxs, B_idxs, 120)
    assert batch.all_observation.shape == (125, 64, 4, 84, 84)
    times = (torch.as_tensor(T_idxs, device="cuda")[None, :] + torch.arange(125, device="cuda")[:, None]) % buf.T
    bi = torch.as_tensor(B_idxs, device="cuda")[None, :].expand_as(times)
    for c in range(4):
        assert torch.equal(batch.all_observation[:, :, c], buf.samples_frames[times + c, bi])
    assert torch.equal(batch.return_, buf.samples_return_[times[:120], bi[:120]])
