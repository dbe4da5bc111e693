"""Oracle: Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) in numpy - the
generator behind the device-side ``Categorical.sample`` (rlpyt_b200/csrc/categorical.cu).  The reference draws
actions with ``torch.multinomial`` (rlpyt/distributions/categorical.py:25-30), whose CPU and CUDA generators differ
from each other, so draw-for-draw parity with the reference does not exist; what is pinned instead is (i) the
inverse-CDF rule (oracle/pg_loss.py:sample_categorical, injected uniforms) and (ii) the uniform stream itself,
restated here.  Test infrastructure only.

Counter = (row_lo, row_hi, call_lo, call_hi), key = (seed_lo, seed_hi); the first output word, top 24 bits,
scaled to [0, 1)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10_word0(rows, call, seed):
    rows = np.asarray(rows, dtype=np.uint64)
    c0, c1 = rows & MASK, rows >> np.uint64(32)
    c2 = np.full_like(rows, np.uint64(call) & MASK)
    c3 = np.full_like(rows, np.uint64(call) >> np.uint64(32))
    k0, k1 = int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2                       # 32 x 32 -> 64 bit products
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32)


def uniforms(n_rows, call, seed):
    """float32 uniforms in [0, 1) for rows 0..n_rows-1 of draw number ``call`` under ``seed``."""
    w = philox4x32_10_word0(np.arange(n_rows), call, seed)
    return ((w >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)
