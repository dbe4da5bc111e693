"""Pin oracle/sum_tree.py + oracle/replay.py against the reference's replay stack (bit-exact:
sampled indices, priorities, IS weights, gathered frames with done-blanking, n-step returns,
the whole fp64 tree) and the known answers of SURVEY.md 9.3 / 9.4."""
import numpy as np
import pytest

from oracle.replay import FrameReplay
from oracle.sum_tree import SumTree
from replay_cases import CASES, drive, case_config


def make(c):
    return FrameReplay(c["obs_shape"], c["size"], c["B"], discount=c["discount"], n_step_return=c["n_step"],
                       prioritized=c["prioritized"], alpha=0.6, beta=0.4, default_priority=1, unique=c["unique"])


@pytest.mark.parametrize("name", CASES)
def test_replay_stream_bit_exact(golden, name):
    g = golden("replay")
    buf = drive(g, name, make, check_root=lambda b: b.tree.tree[0])
    if buf.prioritized:
        assert np.array_equal(buf.tree.tree, g[f"{name}/final_tree"])
    want = g[f"{name}/final_frames"]
    assert np.array_equal(buf.frames[:len(want)] if len(want) != len(buf.frames) else buf.frames, want)
    assert np.array_equal(buf.return_, g[f"{name}/final_return"])
    assert np.array_equal(buf.done_n, g[f"{name}/final_done_n"])


def test_sum_tree_known_answers_survey_9_3(golden):
    g = golden("replay")
    tree = SumTree(T=6, B=2, off_backward=2, off_forward=1, default_value=1.0)
    assert (tree.tree_levels, len(tree.tree), tree.low_idx, tree.high_idx) == (5, 31, 15, 27)
    roots = []
    for k in range(5):
        tree.advance(2)
        assert np.array_equal(tree.tree, g[f"tree_kat/adv{k}"])
        roots.append(tree.tree[0])
    assert roots == [0, 2, 6, 6, 6]
    assert np.array_equal(tree.priorities.T, [[1, 1, 0, 0, 0, 1], [1, 1, 0, 0, 0, 1]])
    np.random.seed(3)
    (T_idxs, B_idxs), pri = tree.sample(5)
    assert list(T_idxs) == [1, 5, 0, 1, 5] and list(B_idxs) == [1, 0, 1, 1, 1] and list(pri) == [1] * 5
    tree.update_batch_priorities(np.array([0.5, 2.0, 3.0, 0.25, 4.0]))
    assert np.array_equal(tree.tree, g["tree_kat/after_update"])
    assert np.array_equal(tree.priorities.T, [[1, 1, 0, 0, 0, 2], [3, 0.5, 0, 0, 0, 4]]) and tree.tree[0] == 11.5
    idx, _ = tree.find(np.array([0, 0.1, 0.5, 0.999999, 1.0]))
    assert list(idx) == [15, 16, 25, 26, 26]
    for lvl in range(tree.tree_levels):  # every level sums to the root
        assert tree.tree[2 ** lvl - 1: 2 ** (lvl + 1) - 1].sum() == 11.5


def test_frame_replay_known_answers_survey_9_4():
    """size=16,B=2,n_frames=3,H=W=1,discount=.5,n_step=2: frame id = global step+1 on b=0."""
    buf = FrameReplay((3, 1, 1), 16, 2, discount=0.5, n_step_return=2)
    assert (buf.T, buf.off_backward, buf.off_forward, buf.frames.shape) == (8, 2, 2, (10, 2, 1, 1))
    step = 0
    hist = [0, 0]

    def batch(T):
        nonlocal step, hist
        obs = np.zeros((T, 2, 3, 1, 1), np.uint8)
        act, rew, done = np.zeros((T, 2), np.int64), np.zeros((T, 2), np.float32), np.zeros((T, 2), bool)
        for t in range(T):
            hist = hist[1:] + [step + 1] if len(hist) == 3 else hist + [step + 1]
            hist = hist[-3:]
            obs[t, 0, :, 0, 0] = hist
            act[t, 0], rew[t, 0], done[t, 0] = step % 4, step + 1, step == 2
            step += 1
        return dict(observation=obs, action=act, reward=rew, done=done)
    buf.append_samples(batch(4))
    assert list(buf.frames[:, 0, 0, 0]) == [0, 0, 1, 2, 3, 4, 0, 0, 0, 0]
    assert list(buf.return_[:, 0]) == [2, 3.5, 3, 0, 0, 0, 0, 0.5] and list(buf.done_n[:, 0]) == [0, 1, 1, 0, 0, 0, 0, 0]
    assert buf.tree.tree[0] == 0
    buf.append_samples(batch(4))
    assert list(buf.frames[:, 0, 0, 0]) == [7, 8, 1, 2, 3, 4, 5, 6, 7, 8]
    assert list(buf.return_[:, 0]) == [2, 3.5, 3, 6.5, 8, 9.5, 11, 0.5]
    b = buf.extract_batch(np.array([1, 2, 3, 4, 5]), np.zeros(5, np.int64))
    assert b["observation"][:, :, 0, 0].tolist() == [[8, 1, 2], [1, 2, 3], [0, 0, 4], [0, 4, 5], [4, 5, 6]]
    assert list(b["prev_action"]) == [0, 1, 0, 3, 0] and list(b["prev_reward"]) == [1, 2, 0, 4, 5]
    assert list(b["action"]) == [1, 2, 3, 0, 1] and list(b["return_"]) == [3.5, 3, 6.5, 8, 9.5]
    assert list(b["done"]) == [0, 1, 0, 0, 0] and list(b["done_n"]) == [1, 1, 0, 0, 0]
    assert b["target_observation"][:, :, 0, 0].tolist() == [[0, 0, 4], [0, 4, 5], [4, 5, 6], [5, 6, 7], [6, 7, 8]]
    buf.append_samples(batch(4))
    assert buf.t == 4 and list(buf.frames[:, 0, 0, 0]) == [7, 8, 9, 10, 11, 12, 5, 6, 7, 8]
