"""Eval scorer: all-pairs cosine scores of user factors against song factors with a fused top-k
(generalises DCUE.predict / model.sim, dcrecommend/nn/dcue.py:495-513; BASELINE cfg5)."""
from __future__ import annotations

import torch

import os

from . import _lib as L

COS_EPS = 1e-8


def _roundup(a, b):
    return (a + b - 1) // b * b


def normalize_factors(x, fmt=L.FMT_F16):
    """[rows, F] fp32 -> row-normalised 16-bit K-major operand panels (device buffer, Kp)."""
    if not x.is_cuda:
        raise RuntimeError("factors must be CUDA tensors: the DCUE B200 path has no CPU fallback")
    x = x.contiguous().float()
    rows, F = x.shape
    Kp = _roundup(F, 16)
    if Kp > 128:
        raise NotImplementedError("feature_dim > 128 is not supported by the top-k scorer")
    out = torch.zeros((Kp // 8) * _roundup(max(rows, 1), 128) * 8, dtype=torch.int16, device=x.device)
    L.call("dcue_normalize_rows", x.data_ptr(), rows, F, COS_EPS, Kp, fmt, out.data_ptr(), L.stream())
    return out, Kp


def topk_scores(user_factors, item_factors, k, item_offset=0, normalized_items=None):
    """-> (scores [U,k] fp32 descending, idx [U,k] int64 = song row + item_offset; -1 = missing).
    `normalized_items` lets a caller reuse the packed song factors across user batches."""
    n_users, n_items = user_factors.shape[0], item_factors.shape[0]
    dev = user_factors.device
    un, Kp = normalize_factors(user_factors)
    inn, Kp2 = normalized_items if normalized_items is not None else normalize_factors(item_factors)
    assert Kp == Kp2
    scores = torch.empty(n_users, k, dtype=torch.float32, device=dev)
    idx = torch.empty(n_users, k, dtype=torch.int64, device=dev)
    if os.environ.get("DCUE_TOPK_2PASS", "1") != "0":
        # threshold pre-pass on a song sample + full pass from the seeded thresholds; the few users whose seed was too
        # high are re-scored exactly below, so the result is the same top-k
        nws = L.query("dcue_topk_2pass_ws_bytes", L.IMPL_TC, n_users, n_items, k)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        n_failed = torch.zeros(1, dtype=torch.int32, device=dev)
        L.call("dcue_topk_scores_2pass", L.IMPL_TC, un.data_ptr(), n_users, inn.data_ptr(), n_items, Kp, L.FMT_F16, k, item_offset,
               scores.data_ptr(), idx.data_ptr(), n_failed.data_ptr(), ws.data_ptr(), nws, L.stream())
        if int(n_failed.item()):
            rows = torch.nonzero(idx[:, 0] == -2).flatten()
            s2, i2 = _topk_exact(user_factors[rows], inn, Kp, n_items, k, item_offset)
            scores[rows], idx[rows] = s2, i2
        return scores, idx
    nws = L.query("dcue_topk_ws_bytes", L.IMPL_TC, n_users, n_items, k)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    L.call("dcue_topk_scores", L.IMPL_TC, un.data_ptr(), n_users, inn.data_ptr(), n_items, Kp, L.FMT_F16, k, item_offset,
           scores.data_ptr(), idx.data_ptr(), ws.data_ptr(), nws, L.stream())
    return scores, idx


def sample_scores(users_packed, n_users, items_packed, n_items, k):
    """Global-threshold protocol, step 1 (csrc/topk.cu dcue_topk_sample): per user the r best scores of this shard's song
    sample, descending -> [n_users, r] (None when the stream is too short to sample)."""
    (un, Kp), (inn, Kp2) = users_packed, items_packed
    assert Kp == Kp2
    r = int(L.lib().dcue_topk_sample_r(n_users, n_items, k))
    if r == 0:
        return None
    out = torch.empty(n_users, r, dtype=torch.float32, device=un.device)
    nws = L.query("dcue_topk_sample_ws_bytes", n_users, n_items, k)
    ws = torch.empty(nws, dtype=torch.uint8, device=un.device)
    L.call("dcue_topk_sample", L.IMPL_TC, un.data_ptr(), n_users, inn.data_ptr(), n_items, Kp, L.FMT_F16, k, out.data_ptr(),
           ws.data_ptr(), nws, L.stream())
    return out


def topk_scores_seeded(users_packed, n_users, items_packed, n_items, k, thr, item_offset=0):
    """Step 2: every song of this shard above the user's threshold (at most the k best), descending; short lists are padded
    with (-inf, -1)."""
    (un, Kp), (inn, _) = users_packed, items_packed
    dev = un.device
    scores = torch.empty(n_users, k, dtype=torch.float32, device=dev)
    idx = torch.empty(n_users, k, dtype=torch.int64, device=dev)
    nws = L.query("dcue_topk_seeded_ws_bytes", n_users)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    L.call("dcue_topk_scores_seeded", L.IMPL_TC, un.data_ptr(), n_users, inn.data_ptr(), n_items, Kp, L.FMT_F16, k, item_offset,
           thr.contiguous().data_ptr(), scores.data_ptr(), idx.data_ptr(), ws.data_ptr(), nws, L.stream())
    return scores, idx


def _topk_exact(user_factors, inn, Kp, n_items, k, item_offset):
    """single-pass scorer (thresholds start at -inf) for a few users."""
    n_users = user_factors.shape[0]
    dev = user_factors.device
    un, _ = normalize_factors(user_factors)
    scores = torch.empty(n_users, k, dtype=torch.float32, device=dev)
    idx = torch.empty(n_users, k, dtype=torch.int64, device=dev)
    nws = L.query("dcue_topk_ws_bytes", L.IMPL_TC, n_users, n_items, k)
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    L.call("dcue_topk_scores", L.IMPL_TC, un.data_ptr(), n_users, inn.data_ptr(), n_items, Kp, L.FMT_F16, k, item_offset,
           scores.data_ptr(), idx.data_ptr(), ws.data_ptr(), nws, L.stream())
    return scores, idx


def merge_topk_parts(s, i):
    """[parts, U, k] stacked per-shard lists (each descending) -> merged (scores [U,k], idx [U,k]); no staging copy."""
    s, i = s.contiguous(), i.contiguous()
    parts, n_users, k = s.shape
    out_s = torch.empty(n_users, k, dtype=torch.float32, device=s.device)
    out_i = torch.empty(n_users, k, dtype=torch.int64, device=s.device)
    if n_users:
        L.call("dcue_topk_merge", s.data_ptr(), i.data_ptr(), parts, n_users, k, out_s.data_ptr(), out_i.data_ptr(), L.stream())
    return out_s, out_i


def merge_topk(scores_parts, idx_parts):
    """Merge per-shard top-k lists ([parts][U,k], each descending) into the global top-k
    (song-sharded eval: one part per GPU)."""
    parts = len(scores_parts)
    s = torch.stack([t.contiguous() for t in scores_parts]).contiguous()
    i = torch.stack([t.contiguous() for t in idx_parts]).contiguous()
    _, n_users, k = s.shape
    out_s = torch.empty(n_users, k, dtype=torch.float32, device=s.device)
    out_i = torch.empty(n_users, k, dtype=torch.int64, device=s.device)
    L.call("dcue_topk_merge", s.data_ptr(), i.data_ptr(), parts, n_users, k, out_s.data_ptr(), out_i.data_ptr(), L.stream())
    return out_s, out_i
