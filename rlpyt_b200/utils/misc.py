"""Minibatch index iterator (mirror of ``rlpyt/utils/misc.py:6-17``)."""
import numpy as np


def iterate_mb_idxs(data_length, minibatch_size, shuffle=False):
    """Yield index batches; the shuffle draws from the GLOBAL numpy MT19937 stream exactly like
    the reference (one ``np.random.shuffle(arange(n))`` per call), so with equal seeds the
    minibatch composition is identical; the trailing ``data_length % minibatch_size`` indices
    are dropped (misc.py:13)."""
    order = None
    if shuffle:
        order = np.arange(data_length)
        np.random.shuffle(order)
    for start in range(0, data_length - minibatch_size + 1, minibatch_size):
        sl = slice(start, start + minibatch_size)
        yield order[sl] if shuffle else sl
