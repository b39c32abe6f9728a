"""The counter-based generator behind the in-kernel dropout masks: the numpy restatement
(oracle/philox.py) against the Random123 known-answer vectors of Philox4x32-10, and the statistics
of the keep masks built from it."""
import numpy as np

from oracle import philox


def _kat(ctr, key):
    out = philox.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
    return [int(v) for v in out]


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert _kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_keep_masks_are_bernoulli_and_counter_based():
    keep = philox.bst_keep_masks(seed=1234567, offset=3, n_rows=4096, p=0.1)
    assert keep.shape == (3, 4096, 16)
    n = keep.size
    assert abs((1.0 - keep.mean()) - 6554 / 65536) < 4 * (0.1 * 0.9 / n) ** 0.5
    again = philox.bst_keep_masks(seed=1234567, offset=3, n_rows=4096, p=0.1)
    other = philox.bst_keep_masks(seed=1234567, offset=4, n_rows=4096, p=0.1)
    assert np.array_equal(keep, again) and not np.array_equal(keep, other)
    assert not np.array_equal(keep[0], keep[1])                 # the sites draw different bits
    assert philox.bst_keep_masks(1, 0, 8, 0.0).all()            # p = 0 keeps everything
    prefix = philox.bst_keep_masks(seed=1234567, offset=3, n_rows=100, p=0.1)
    assert np.array_equal(prefix, keep[:, :100])                # a row's bits do not depend on the batch
