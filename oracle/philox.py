"""TEST INFRASTRUCTURE — numpy restatement of the dropout keep-bits of librank_b200's BST block.

The reference draws its dropout masks from torch's generators (BST/bst.py:57,62,86,90:
nn.Dropout inside BSTTransformer); those draws cannot be replayed inside a fused kernel, so the
kernels derive the masks from a counter-based generator instead (csrc/common.cuh:
philox4x32_10 / dropout_keep16).  This file restates that construction so the tests can hand the
*same* masks to the oracle block (oracle/interactions.py::bst_transformer_block) and compare
outputs and gradients at the fp32 bar.  Philox4x32-10 is Salmon et al., "Parallel random numbers:
as easy as 1, 2, 3" (SC'11); tests/test_philox.py pins it to the Random123 known-answer vectors.
Only tests/ may import this module.
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 [..., 4], key: uint32 [..., 2] -> uint32 [..., 4]."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for _ in range(10):
        p0 = _M0 * c[0]
        p1 = _M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def bst_keep_masks(seed, offset, n_rows, p):
    """Keep masks of the three dropout sites of one block forward: bool [3, n_rows, 16]
    (site 0: w_o output, 1: inside the FFN, 2: FFN output; row = b*T + t)."""
    thr = int(np.rint(np.float32(p) * np.float32(65536.0)))
    rows = np.arange(n_rows, dtype=np.uint64)
    key = np.empty((n_rows, 2), dtype=np.uint32)
    key[:, 0] = seed & 0xFFFFFFFF
    key[:, 1] = ((seed >> 32) & 0xFFFFFFFF) ^ ((offset >> 32) & 0xFFFFFFFF)
    keep = np.empty((3, n_rows, 16), dtype=bool)
    for site in range(3):
        for half in range(2):
            ctr = np.empty((n_rows, 4), dtype=np.uint32)
            ctr[:, 0] = (rows & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            ctr[:, 1] = (rows >> np.uint64(32)).astype(np.uint32)
            ctr[:, 2] = 2 * site + half
            ctr[:, 3] = offset & 0xFFFFFFFF
            w = philox4x32_10(ctr, key)
            for j in range(4):
                keep[site, :, 8 * half + 2 * j] = (w[:, j] & 0xFFFF) >= thr
                keep[site, :, 8 * half + 2 * j + 1] = (w[:, j] >> 16) >= thr
    return keep
