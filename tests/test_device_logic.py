"""CPU restatements of pieces of device-side logic, checked against numpy (round-2 additions: sections 3 and 4 below).  (1) The selection logic of the top-k scorer (csrc/topk.cu):
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


# ----------------------------------------------------------------------------------------------------------------------
# (3) round 2: the cp.async tile loader of the dense-layer GEMM (csrc/linear.cu TileLoader) computes each thread's chunk
# address / byte count once and derives the per-k-chunk values from (r0, rem = rend - r0) alone; (4) the exchanged-row
# reduction of the data-parallel table gradient (csrc/peer.cu peer_mark_rows_kernel + peer_scatter_add_rows_kernel) marks
# the first entry and the entry count of every row with integer atomics, and only first entries sum (in entry order).
TKC, CP_THREADS = 32, 128


def _bytes_reference(rc, tid, it, i0, I, r0, rend):
    """The first version's per-chunk arithmetic (cp_load_tile): -> (bytes, element offset of the source or None)."""
    c = tid + it * CP_THREADS
    if rc:
        li, lr = c // (TKC // 4), (c % (TKC // 4)) * 4
        gi, gr = i0 + li, r0 + lr
        nb = min(16, (rend - gr) * 4) if (gi < I and gr < rend) else 0
        return nb, ((gi, gr) if nb else None)
    lr, li = c >> 4, (c & 15) * 4
    gi, gr = i0 + li, r0 + lr
    nb = min(16, (I - gi) * 4) if (gr < rend and gi < I) else 0
    return nb, ((gi, gr) if nb else None)


def _bytes_hoisted(rc, tid, it, i0, I, r0, rend):
    """TileLoader.init + TileLoader.issue."""
    c = tid + it * CP_THREADS
    rem = rend - r0
    if rc:
        li, lr = c // (TKC // 4), (c % (TKC // 4)) * 4
        gi = i0 + li
        lim = lr if gi < I else (1 << 29)
        nb = max(0, min(16, (rem - lim) * 4))
        return nb, ((gi, lr + r0) if nb else None)
    lr, li = c >> 4, (c & 15) * 4
    gi = i0 + li
    fixed = max(0, min(16, (I - gi) * 4))
    lr_issue = (tid >> 4) + it * (CP_THREADS // 16)
    assert lr_issue == lr
    nb = fixed if lr_issue < rem else 0
    return nb, ((gi, lr + r0) if nb else None)


def test_gemm_tile_loader_hoisting_is_equivalent():
    rng = np.random.default_rng(0)
    cases = [(0, 64, 0, 32), (0, 100, 96, 100), (64, 100, 96, 100), (64, 100, 64, 97), (0, 3, 0, 5), (128, 300, 288, 300)]
    cases += [(int(64 * rng.integers(0, 6)), int(rng.integers(1, 400)), int(32 * rng.integers(0, 10)), int(rng.integers(1, 330)))
              for _ in range(60)]
    for i0, I, r0, rend in cases:
        if r0 >= rend:
            continue
        for rc in (True, False):
            for tid in range(CP_THREADS):
                for it in range(4):
                    assert _bytes_reference(rc, tid, it, i0, I, r0, rend) == _bytes_hoisted(rc, tid, it, i0, I, r0, rend), \
                        (rc, tid, it, i0, I, r0, rend)


def _marked_row_reduction(all_idx, rows, lo, hi):
    """peer_mark_rows_kernel + peer_scatter_add_rows_kernel on numpy arrays (float32 adds in entry order)."""
    n = len(all_idx)
    first = np.zeros(hi - lo, dtype=np.int64)       # n - j of the first entry (atomicMax), 0 = no entry
    cnt = np.zeros(hi - lo, dtype=np.int64)
    for j in np.random.default_rng(1).permutation(n):        # any order: max / count do not depend on it
        r = all_idx[j]
        if lo <= r < hi:
            first[r - lo] = max(first[r - lo], n - j)
            cnt[r - lo] += 1
    out = np.zeros((hi - lo, rows.shape[1]), dtype=np.float32)
    for j in range(n):
        r = all_idx[j]
        if not (lo <= r < hi) or n - first[r - lo] != j:
            continue                                           # an earlier entry owns this row
        acc = np.zeros(rows.shape[1], dtype=np.float32)
        found = 0
        for jj in range(j, n):                                 # the kernel scans 32 entries per ballot and stops at `cnt`
            if all_idx[jj] == r:
                acc = acc + rows[jj]
                found += 1
                if found == cnt[r - lo]:
                    break
        out[r - lo] = acc
    return out


def test_marked_row_reduction_equals_ordered_scatter_add():
    rng = np.random.default_rng(3)
    for n, U, lo, hi in ((64, 10, 0, 10), (2048, 500, 0, 500), (512, 40, 10, 30), (300, 1000, 0, 1000)):
        idx = rng.integers(0, U, n)
        idx[1] = idx[0]
        idx[n - 1] = idx[0]                                    # a row with entries at both ends
        rows = rng.standard_normal((n, 12)).astype(np.float32)
        ref = np.zeros((hi - lo, 12), dtype=np.float32)
        for j in range(n):                                     # sequential fp32 adds in entry order = the kernel's order
            if lo <= idx[j] < hi:
                ref[idx[j] - lo] = ref[idx[j] - lo] + rows[j]
        assert np.array_equal(_marked_row_reduction(idx, rows, lo, hi), ref)
