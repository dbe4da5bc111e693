"""GPU parity of the replay path through the public classes (device sum-tree, frame replay,
fused extraction): the recorded reference streams of tests/golden/replay.npz must be reproduced
BIT-EXACTLY - sampled indices, tree contents (all 2^L-1 fp64 nodes), gathered/blanked frames,
n-step returns, every scalar field; importance weights (fp32 from an fp64 pow) to 1e-6."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from collections import namedtuple  # noqa: E402

from replay_cases import CASES, FIELDS, case_config, drive  # noqa: E402

Example = None


class Adapter:
    """Presents rlpyt_b200's reference-shaped replay classes through the oracle's dict interface."""

    def __init__(self, c):
        from rlpyt_b200.replays.non_sequence.frame import PrioritizedReplayFrameBuffer, UniformReplayFrameBuffer
        from rlpyt_b200.utils.collections import namedarraytuple
        global Example
        Example = namedarraytuple("SamplesToBuffer", ["observation", "action", "reward", "done"])
        ex = Example(observation=np.zeros(c["obs_shape"], np.uint8), action=np.int64(0), reward=np.float32(0),
                     done=np.bool_(False))
        kw = dict(example=ex, size=c["size"], B=c["B"], discount=c["discount"], n_step_return=c["n_step"])
        self.buf = (PrioritizedReplayFrameBuffer(alpha=0.6, beta=0.4, default_priority=1, unique=c["unique"],
                                                 pow_on_host=c.get("pow_on_host", True), **kw)
                    if c["prioritized"] else UniformReplayFrameBuffer(**kw))
        self.c = c

    @property
    def t(self):
        return self.buf.t

    def append_samples(self, s):
        self.buf.append_samples(Example(**{k: torch.from_numpy(v).cuda() for k, v in s.items()}))

    def sample_batch(self, n, random_values=None):
        if self.c["prioritized"]:
            b = self.buf.sample_batch(n, random_values=random_values)
            T_idxs, B_idxs = self.buf._last_idxs
        else:
            st = np.random.get_state()
            T_idxs, B_idxs = self.buf.sample_idxs(n)
            np.random.set_state(st)
            b = self.buf.sample_batch(n)
        out = dict(observation=b.agent_inputs.observation, prev_action=b.agent_inputs.prev_action,
                   prev_reward=b.agent_inputs.prev_reward, action=b.action, return_=b.return_, done=b.done,
                   done_n=b.done_n, target_observation=b.target_inputs.observation,
                   target_prev_action=b.target_inputs.prev_action, target_prev_reward=b.target_inputs.prev_reward,
                   T_idxs=T_idxs, B_idxs=B_idxs)
        if self.c["prioritized"]:
            out["is_weights"] = b.is_weights
        return out

    def update_batch_priorities(self, p):
        self.buf.update_batch_priorities(torch.from_numpy(np.asarray(p)).cuda())


def to_np(x):
    return x.cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class _IsW:
    """is_weights comparison is tolerance-based: patch array_equal for that one field."""


@pytest.mark.parametrize("name", CASES)
def test_replay_stream_bit_exact_on_gpu(golden, name, monkeypatch):
    g = golden("replay")
    import replay_cases
    real_equal = np.array_equal

    def equal(a, b):  # fp32 importance weights: 1e-6 relative; everything else exact
        a, b = np.asarray(a), np.asarray(b)
        if a.dtype == np.float32 and a.ndim == 1 and a.shape == b.shape and a.size and (a.max() == 1.0 == b.max()):
            return np.allclose(a, b, rtol=1e-6, atol=0)
        return real_equal(a, b)
    monkeypatch.setattr(replay_cases.np, "array_equal", equal)
    ad = drive(g, name, Adapter, to_np=to_np, check_root=lambda a: float(a.buf.priority_tree.tree[0].item()))
    monkeypatch.undo()
    buf = ad.buf
    if ad.c["prioritized"]:
        assert np.array_equal(buf.priority_tree.tree.cpu().numpy(), g[f"{name}/final_tree"])
    want = g[f"{name}/final_frames"]
    assert np.array_equal(buf.samples_frames.cpu().numpy()[:len(want)], want)
    assert np.array_equal(buf.samples_return_.cpu().numpy(), g[f"{name}/final_return"])
    assert np.array_equal(buf.samples_done_n.cpu().numpy(), g[f"{name}/final_done_n"])


def test_sum_tree_known_answers_on_gpu(golden):
    from rlpyt_b200.replays.sum_tree import SumTree
    g = golden("replay")
    tree = SumTree(T=6, B=2, off_backward=2, off_forward=1, default_value=1.0)
    for k in range(5):
        tree.advance(2)
        assert np.array_equal(tree.tree.cpu().numpy(), g[f"tree_kat/adv{k}"]), k
    np.random.seed(3)
    (T_idxs, B_idxs), pri = tree.sample(5)
    assert T_idxs.tolist() == [1, 5, 0, 1, 5] and B_idxs.tolist() == [1, 0, 1, 1, 1] and pri.tolist() == [1.0] * 5
    tree.update_batch_priorities(torch.tensor([0.5, 2.0, 3.0, 0.25, 4.0], dtype=torch.float64))
    assert np.array_equal(tree.tree.cpu().numpy(), g["tree_kat/after_update"])
    idx, _ = tree.find(np.array([0, 0.1, 0.5, 0.999999, 1.0]))
    assert idx.tolist() == [15, 16, 25, 26, 26]


@pytest.mark.parametrize("batch_kernel", [True, False])
@pytest.mark.parametrize("T,B,adv", [(3907, 256, 128), (64, 8, 50), (1000, 3, 17)])
def test_sum_tree_random_stream_vs_oracle(T, B, adv, batch_kernel):
    """Config-4 scale (1 M leaves, 21 levels): advance / sample(512) / update loops against the numpy
    oracle - whole tree bit-identical after every operation, sampled indices identical.  ``batch_kernel``: the
    two-launch update (device sort in one CTA) or the generic sorted-segment path (torch.sort + set-leaves)."""
    from oracle.sum_tree import SumTree as Oracle
    from rlpyt_b200.replays.sum_tree import SumTree
    o = Oracle(T, B, off_backward=3, off_forward=3, default_value=1.0)
    d = SumTree(T, B, off_backward=3, off_forward=3, default_value=1.0)
    if not batch_kernel:
        d.BATCH_KERNEL_MAX = 0
    rng = np.random.default_rng(T)
    n_adv = 3 * T // adv + 3 if T < 2000 else 34
    for it in range(n_adv):
        o.advance(adv)
        d.advance(adv)
        if o.tree[0] > 0:
            for _ in range(2):
                u = rng.random(512)
                (To, Bo), po = o.sample(512, random_values=u)
                (Td, Bd), pd = d.sample(512, random_values=u)
                assert np.array_equal(Td.cpu().numpy(), To) and np.array_equal(Bd.cpu().numpy(), Bo)
                assert np.array_equal(pd.cpu().numpy(), po)
                new = (np.abs(rng.standard_normal(512)).astype(np.float32) + 1e-3) ** np.float32(0.6)
                o.update_batch_priorities(new)
                d.update_batch_priorities(torch.from_numpy(new).cuda())
        if it % 5 == 0 or it == n_adv - 1:
            assert np.array_equal(d.tree.cpu().numpy(), o.tree), it
    assert d.t == o.t


@pytest.mark.parametrize("n", [1, 31, 512, 2048])
def test_update_batch_with_folded_pow_equals_pow_then_update(n):
    """``update_batch_priorities(p, alpha)`` (pow evaluated inside the batch kernel) leaves the tree bit-identical to
    the separate pow kernel followed by the generic update - duplicates, unsorted indices and all batch sizes up to
    the kernel's limit included."""
    from rlpyt_b200.replays.sum_tree import SumTree, _pow_alpha_f64
    trees = [SumTree(300, 7, off_backward=2, off_forward=2, default_value=1.0) for _ in range(2)]
    trees[1].BATCH_KERNEL_MAX = 0
    rng = np.random.default_rng(n)
    for t in trees:
        t.advance(250)
    for rep in range(3):
        u = rng.random(n)
        pri = torch.from_numpy((np.abs(rng.standard_normal(n)) * 3 + 1e-4).astype(np.float32)).cuda()
        for t in trees:
            t.sample(n, random_values=u)
        trees[0].update_batch_priorities(pri, alpha=0.6)
        trees[1].update_batch_priorities(_pow_alpha_f64(pri, 0.6, pri.device))
        assert torch.equal(trees[0].tree, trees[1].tree), rep
        nodes = trees[0].tree.cpu().numpy()
        np.testing.assert_allclose(nodes[0], nodes[trees[0].low_idx:trees[0].high_idx].sum(), rtol=1e-12)


def test_pow_alpha_kernel_is_correctly_rounded_and_within_1ulp_of_numpy():
    """The device ``priorities ** alpha`` returns the correctly rounded float32 power (computed
    through fp64); numpy's SVML float32 ``power`` - what the reference runs - is within 1 ulp of it."""
    from rlpyt_b200.replays.non_sequence.prioritized import PrioritizedReplay

    class P(PrioritizedReplay):
        def __init__(self):
            self.alpha, self.device, self.pow_on_host = 0.6, torch.device("cuda"), False
    x = (np.abs(np.random.default_rng(0).standard_normal(100000)) * 3 + 1e-4).astype(np.float32)
    got = P()._pow_alpha(torch.from_numpy(x)).cpu().numpy()
    exact = (x.astype(np.float64) ** np.float64(np.float32(0.6))).astype(np.float32)
    assert np.array_equal(got, exact.astype(np.float64))
    ref = x ** 0.6   # numpy: float32 array ** python float -> float32 pow
    assert ref.dtype == np.float32
    ulp = np.abs(got.astype(np.float32).view(np.int32) - ref.view(np.int32))
    assert ulp.max() <= 1


def test_replay_stream_device_pow_statistics(golden):
    """Same recorded stream with the default device pow: indices/frames are bit-exact while the tree
    is identical (first update not yet applied) and stay within 1e-6 relative afterwards."""
    g = golden("replay")
    c = case_config(g, "mid_pri")
    c["pow_on_host"] = False
    ad = Adapter(c)
    import replay_cases
    np.random.seed(c["seed"])
    for i, s in enumerate(replay_cases.replay_stream(c["seed"], c["n_batches"], c["batch_T"], c["B"], c["obs_shape"], 4, 0.1)):
        ad.append_samples(s)
        if f"mid_pri/b{i}/uniforms" in g.files:
            b = ad.sample_batch(c["batch_B"], random_values=g[f"mid_pri/b{i}/uniforms"])
            ad.update_batch_priorities(g[f"mid_pri/b{i}/new_pri"])
    np.testing.assert_allclose(ad.buf.priority_tree.tree.cpu().numpy(), g["mid_pri/final_tree"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("vector", [False, True])
def test_extract_full_config_size(vector, monkeypatch):
    """1M-frame-shaped extraction at batch 512 (84x84, 4 frames, n-step 3) on a smaller ring: frames
    bit-identical to the oracle's gather, blanking included - through the bulk-copy kernel (default) and the
    vector-load fallback (csrc/replay.cu)."""
    from oracle.replay import FrameReplay
    if vector:
        monkeypatch.setenv("RLPYT_B200_REPLAY_VECTOR", "1")
    c = dict(obs_shape=(4, 84, 84), size=64 * 40, B=40, discount=0.99, n_step=3, prioritized=True, unique=False)
    ad = Adapter(c)
    o = FrameReplay(c["obs_shape"], c["size"], c["B"], discount=0.99, n_step_return=3)
    import replay_cases
    for s in replay_cases.replay_stream(5, 6, 16, 40, c["obs_shape"], 6, 0.05):
        ad.append_samples(s)
        o.append_samples(s)
    u = np.random.default_rng(1).random(512)
    want = o.sample_batch(512, random_values=u)
    got = ad.sample_batch(512, random_values=u)
    for k in FIELDS:
        assert np.array_equal(to_np(got[k]), want[k]), k
    assert np.array_equal(to_np(got["T_idxs"]), want["T_idxs"])
