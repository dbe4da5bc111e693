"""Pin oracle/ppo.py + oracle/atari_ff.py against two iterations of the reference's PPO (and its
normalised / valid-masked variant) on a tiny AtariFf problem (tests/golden/ppo.npz): same OptInfo
rows and same updated weights.  CPU conv/GEMM reductions may be re-ordered between runs of
different thread counts, so this is held to 1e-5 relative rather than bit equality."""
import numpy as np
import pytest
import torch

from oracle import atari_ff
from oracle.ppo import PpoOracle

T, B, IMAGE, A = 8, 6, (4, 36, 36), 5
CFG = {
    "ppo": dict(gae_lambda=0.98, minibatches=2, epochs=2, mid_batch_reset=True),
    "ppo_valid_norm": dict(gae_lambda=1, minibatches=3, epochs=2, normalize_advantage=True, ratio_clip=0.2,
                           mid_batch_reset=False),
}


def batch(g, name, itr):
    return [g[f"{name}/itr{itr}/{k}"] for k in ("obs", "action", "reward", "done", "value", "old_prob", "bv")]


@pytest.mark.parametrize("name", list(CFG))
def test_ppo_oracle_two_iterations(golden, name):
    g = golden("ppo")
    torch.set_num_threads(1)
    sd0 = atari_ff.init_state_dict(IMAGE, A, seed=int(g[f"{name}/sd0_seed"][0]))
    assert abs(float(sum(v.double().sum() for v in sd0.values())) - float(g[f"{name}/sd0_check"][0])) < 1e-9
    o = PpoOracle(sd0, n_itr=4, **CFG[name])
    np.random.seed(77)
    for itr in range(2):
        info = o.optimize_agent(itr, *batch(g, name, itr))
        for f in ("loss", "gradNorm", "entropy", "perplexity"):
            np.testing.assert_allclose(info[f], g[f"{name}/itr{itr}/opt_{f}"], rtol=1e-5, atol=1e-7, err_msg=f)
        sd = o.state_dict()
        for k, v in sd.items():
            want = g[f"{name}/itr{itr}/sd/{k}"]
            got = v.numpy()[:8] if k == "conv.head.model.0.weight" else v.numpy()
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-6, err_msg=k)
