"""Host restatements of small integer schemes the kernels rely on (CPU; no extension needed).

Each test restates, in numpy / plain Python, a piece of device arithmetic whose correctness argument lives
in a kernel comment, and checks the argument over the whole range the kernel can see:
  * afm_tc.cu::small_div — item / d through a float reciprocal instead of an integer division;
  * din.cu::plan_tile_tc — tile planning from prefix sums, against the sequential loop it replaced;
  * din.cu::kstage_chunk — the 16-byte-chunk swizzle of the row staging buffer is conflict-free;
  * afm_tc.cu recheck — the 64-bit "columns inside the band" mask visits exactly the flagged columns.
"""
import numpy as np

K_ROWS = 128          # din.cu::kRows
K_GROUP = 8           # din.cu::kTcGroup


def test_small_div_is_exact_on_the_kernel_range():
    # afm_tc.cu: items < 1024 (S*F*D/4 <= 512 staged rows, 2..4 loop trips of 128), divisors D/4 in 1..8, F in 2..16
    item = np.arange(1024, dtype=np.int32)
    for d in range(1, 17):
        inv = np.float32(1.0) / np.float32(d)
        q = ((item.astype(np.float32) + np.float32(0.5)) * inv).astype(np.int32)       # cvt.rzi of a positive value
        assert np.array_equal(q, item // d), d


def plan_loop(length, n_samples, s_begin, tid):
    """The sequential planner the prefix-sum version replaced (whole samples while they fit in 128 rows)."""
    rows, s = 0, s_begin
    my_s, my_t, on = s_begin, 0, False
    while s < n_samples and rows + length[s] <= K_ROWS:
        if rows <= tid < rows + length[s]:
            my_s, my_t, on = s, tid - rows, True
        rows += length[s]
        s += 1
    return s, rows, my_s, my_t, on


def plan_prefix(length, n_samples, s_begin, tid):
    """din.cu::plan_tile_tc as shipped: straight-line code over the nine prefix sums."""
    pre = [0] * (K_GROUP + 1)
    for i in range(K_GROUP):
        pre[i + 1] = pre[i] + (length[i] if i < n_samples else 0)
    base = pre[s_begin]
    s_end, n_rows, my_s, start = s_begin, 0, s_begin, 0
    for i in range(K_GROUP):
        end_i = pre[i + 1] - base
        fits = s_begin <= i < n_samples and end_i <= K_ROWS
        if fits:
            s_end, n_rows = i + 1, end_i
        if fits and tid >= end_i:
            my_s, start = i + 1, end_i
    on = tid < n_rows
    return s_end, n_rows, (my_s if on else s_begin), (tid - start if on else 0), on


def test_prefix_sum_tile_planner_matches_the_loop():
    rng = np.random.default_rng(0)
    cases = 0
    for trial in range(400):
        T = int(rng.choice([1, 7, 20, 50, 64, 128]))
        n_samples = int(rng.integers(1, K_GROUP + 1))
        length = [int(x) for x in rng.integers(0, T + 1, K_GROUP)]
        if trial % 5 == 0:
            length = [T] * K_GROUP                      # every history full
        if trial % 7 == 0:
            length[int(rng.integers(0, K_GROUP))] = 0   # an empty history in the middle
        s_begin = 0
        while s_begin < n_samples:                      # walk the group tile by tile, as the kernel does
            ref0 = plan_loop(length, n_samples, s_begin, 0)
            for tid in range(K_ROWS):
                assert plan_prefix(length, n_samples, s_begin, tid) == plan_loop(length, n_samples, s_begin, tid), \
                    (length, n_samples, s_begin, tid)
            assert ref0[0] > s_begin                    # progress: T <= 128, so at least one sample fits
            s_begin = ref0[0]
            cases += 1
    assert cases > 400


def test_row_staging_swizzle_is_conflict_free():
    # kstage rows are 64 bytes apart; a warp instruction moves chunk c (16 bytes) of 32 consecutive rows.
    # Eight threads' 16-byte pieces fill the 32 banks once: the access is conflict-free when every group of
    # eight consecutive rows covers the eight 16-byte bank groups.
    for c in range(4):
        for first in range(0, K_ROWS, 8):
            groups = {((row * 64 + 16 * (c ^ ((row >> 1) & 3))) // 16) % 8 for row in range(first, first + 8)}
            assert len(groups) == 8, (c, first)
    # and it is a permutation of the four chunks of a row (nothing overwritten, everything read back)
    for row in range(K_ROWS):
        assert sorted(c ^ ((row >> 1) & 3) for c in range(4)) == [0, 1, 2, 3]


def test_band_mask_visits_exactly_the_flagged_columns():
    rng = np.random.default_rng(1)
    for _ in range(200):
        x = rng.standard_normal(64).astype(np.float32) * np.float32(1e-3)
        band = np.float32(abs(rng.standard_normal()) * 2e-4)
        near0 = near1 = 0
        for j in range(32):
            if abs(x[j]) < band:
                near0 |= 1 << j
            if abs(x[32 + j]) < band:
                near1 |= 1 << j
        near = (near1 << 32) | near0
        visited = []
        while near:
            low = near & -near                           # __ffsll(near) - 1
            visited.append(low.bit_length() - 1)
            near &= near - 1
        assert visited == [j for j in range(64) if abs(x[j]) < band]
