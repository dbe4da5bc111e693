"""oracle/replay_sequence.py against fixtures recorded from the reference's sequence replay buffers
(tests/golden/make_golden.py: gen_seq_replay - UniformSequenceReplayFrameBuffer /
PrioritizedSequenceReplayFrameBuffer with rnn_state_interval 0, 1 and > 1, input priorities, wrap at both ends)."""
import numpy as np
import pytest

from oracle import replay_sequence as R
from replay_cases import SEQ_CASES as CASES, seq_replay_case as replay_case

def test_extract_sequences_known_answers(golden):
    g = golden("seq_replay")
    out = R.extract_sequences(g["extract_kat/arr"], g["extract_kat/T_idxs"], g["extract_kat/B_idxs"], 4)
    assert np.array_equal(out, g["extract_kat/out"])
    # a negative start puts the wrapped rows at the END (rlpyt/utils/misc.py:49-51): start -1 of column 0 -> rows 0,1,2,9
    assert out[:, 0].tolist() == [0, 3, 6, 27]


@pytest.mark.parametrize("name", CASES)
def test_sequence_replay_stream_bit_exact(golden, name):
    g = golden("seq_replay")

    def make(obs_shape, size, B, rsi, batch_T, discount, n_step, prioritized, input_pri, pri_shift):
        return R.SequenceFrameReplay(obs_shape, size, B, rsi, batch_T, rnn_state_shapes=dict(h=(1, 3), c=(1, 3)),
                                     discount=discount, n_step_return=n_step, prioritized=prioritized, alpha=0.6, beta=0.9,
                                     default_priority=1, input_priorities=input_pri, input_priority_shift=pri_shift)

    buf = replay_case(g, name, make, lambda b, s, p: b.append_samples(s, priorities=p),
                      lambda b, n, u: b.sample_batch(n, random_values=u), lambda b, p: b.update_batch_priorities(p))
    if buf.seq_prioritized:
        assert np.array_equal(buf.tree.tree, g[f"{name}/final_tree"])
    assert np.array_equal(buf.return_, g[f"{name}/final_return"])
    assert np.array_equal(buf.done_n, g[f"{name}/final_done_n"])
    if buf.rsi > 1:
        assert np.array_equal(buf.rnn_state["h"], g[f"{name}/final_rnn_h"])
