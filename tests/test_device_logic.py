"""CPU restatements of two pieces of device-side logic, checked against numpy.  (1) The selection logic of the top-k scorer (csrc/topk.cu):
  * the order-preserving float -> uint32 key map and the bit-by-bit radix select of the k-th largest key
    (select_topk_inplace: common bits skipped, `rem` copies of the threshold key kept),
  * the seeding arithmetic of dcue_topk_scores_2pass (r-th best of every s-th tile => about 4k candidates, and how
    rarely fewer than k).
(2) The word-wide row masks of the fused layer-1 backward (csrc/conv_tc.cu: tc_wgrad_unpool_kernel).
The kernels themselves are checked on the GPU (tests/test_gpu_kernels.py); this pins the algorithm they implement."""
import math

import numpy as np


def f2key(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32)
    return np.where(b >> 31 == 1, ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def key2f(k):
    k = np.asarray(k, dtype=np.uint32)
    return np.where(k >> 31 == 1, k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32).view(np.float32)


def radix_select(keys, k):
    """-> (T, rem): key of the k-th largest and how many entries equal to T belong to the top k."""
    keys = np.asarray(keys, dtype=np.uint32)
    aand, oor = np.bitwise_and.reduce(keys), np.bitwise_or.reduce(keys)
    diff = int(aand ^ oor)
    prefix, decided, rem = int(aand), (~diff) & 0xFFFFFFFF, k
    for bit in range(31, -1, -1):
        b = 1 << bit
        if not diff & b:
            continue
        m, want = decided | b, prefix | b
        c = int(np.count_nonzero((keys & np.uint32(m)) == np.uint32(want)))
        if c >= rem:
            prefix |= b
        else:
            rem -= c
        decided |= b
    return prefix, rem


def test_key_map_preserves_order_and_round_trips():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(1000).astype(np.float32), np.float32([0.0, -0.0, 1e-38, -1e-38, np.inf, -np.inf])])
    k = f2key(x)
    order = np.lexsort((k, x))                                       # by value; the two zeros (equal values) by key
    assert (np.diff(k[order].astype(np.int64)) >= 0).all()          # x ascending -> keys non-decreasing
    xs = x[order]
    assert (np.diff(k[order].astype(np.int64))[np.diff(xs) > 0] > 0).all()   # strictly larger value -> strictly larger key
    assert f2key(np.float32(-0.0)) < f2key(np.float32(0.0))         # the only pair of equal values with distinct keys
    assert (key2f(k).view(np.uint32) == x.view(np.uint32)).all()


def test_radix_select_matches_partition():
    rng = np.random.default_rng(1)
    for n, k, ties in [(512, 100, False), (385, 100, True), (100, 100, False), (300, 1, True), (512, 256, True), (64, 25, False)]:
        x = rng.standard_normal(n).astype(np.float32) * 0.1
        if ties:
            x[rng.integers(0, n, n // 3)] = x[0]                     # many duplicates, possibly at the threshold
            x[rng.integers(0, n, 5)] *= -1
        keys = f2key(x)
        T, rem = radix_select(keys, k)
        kth = np.sort(x)[::-1][k - 1]
        assert key2f(np.uint32(T)) == kth
        n_gt = int(np.count_nonzero(keys > np.uint32(T)))
        n_eq = int(np.count_nonzero(keys == np.uint32(T)))
        assert n_gt + rem == k and 1 <= rem <= n_eq                  # the kernel keeps all > T and `rem` entries == T
    same = f2key(np.full(40, 0.25, dtype=np.float32))                # no varying bit at all
    assert radix_select(same, 7) == (int(same[0]), 7)


def test_seeding_leaves_k_candidates_with_high_probability():
    """Seed = r-th best of a 1/s sample; the number of full-stream scores above it is ~ negative-binomial with mean
    r*s and relative spread 1/sqrt(r).  With the plan's (s, r) fewer than k are left about once per 1e4 users."""
    k = 100
    C = 4 * k
    s = min(16, C // 20)
    r = -(-C // s)
    assert (s, r) == (16, 25)
    rng = np.random.default_rng(2)
    trials, N = 20000, 500000
    # the seed's rank in the full stream: the r-th sampled order statistic of N/s uniform draws -> Beta(r, N/s - r + 1)
    q = rng.beta(r, N // s - r + 1, size=trials)
    above = rng.binomial(N, q)
    assert abs(above.mean() / (r * s) - 1) < 0.05
    assert (above < k).mean() < 1e-3
    # closed form tail of the dominant term (normal approximation): z = (mean - k) / (mean / sqrt(r))
    z = (r * s - k) / (r * s / math.sqrt(r))
    assert z > 3.5


def test_unpool_row_mask_bit_trick():
    """The fused layer-1 backward (conv_tc.cu: tc_wgrad_unpool_kernel) turns the four argmax-code bytes of a word into two
    16-bit-lane masks with word-wide logic: XOR with the row, exact zero-byte test, byte -> 0xff, byte permutes.
    Restated here for every byte value that can occur (codes 0..3 and the 'no window' filler 0xff) and every row."""
    def prmt(x, sel):   # __byte_perm(x, 0, sel) for selector nibbles 0..3
        by = [(x >> (8 * i)) & 0xFF for i in range(4)]
        return sum(by[(sel >> (4 * i)) & 0x7] << (8 * i) for i in range(4))
    vals = [0, 1, 2, 3, 0xFF]
    for j in range(4):
        for b0 in vals:
            for b1 in vals:
                for b2 in vals:
                    for b3 in vals:
                        cw = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24)
                        t = cw ^ (j * 0x01010101)
                        nz = (((t & 0x7F7F7F7F) + 0x7F7F7F7F) | t) & 0xFFFFFFFF
                        m8 = ((((~nz) & 0x80808080) >> 7) * 0xFF) & 0xFFFFFFFF
                        lo, hi = prmt(m8, 0x1100), prmt(m8, 0x3322)
                        want_lo = (0xFFFF if b0 == j else 0) | (0xFFFF0000 if b1 == j else 0)
                        want_hi = (0xFFFF if b2 == j else 0) | (0xFFFF0000 if b3 == j else 0)
                        assert (lo, hi) == (want_lo, want_hi)
