"""Pins oracle/dcue_oracle.py against outputs of the reference itself
(tests/golden/ref_*.pt, produced by oracle/make_golden.py in the build container)."""
import os

import pytest
import torch

from oracle import dcue_oracle as O
from oracle import fixtures

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _close(a, b, tol=2e-5):
    a, b = a.double(), b.double()
    denom = b.abs().max().clamp_min(1e-30)
    assert ((a - b).abs().max() / denom).item() <= tol, ((a - b).abs().max() / denom).item()


@pytest.mark.parametrize("mt", O.MODEL_TYPES)
def test_oracle_matches_reference_train_step(mt):
    g = _load("ref_%s.pt" % mt)
    p = fixtures.make_params(mt, seed=0, user_count=g["U"])
    u, pos, neg = fixtures.make_inputs(g["B"], g["N"], g["U"], seed=1)
    u[1] = u[0]
    assert torch.equal(u, g["u"])
    r = O.train_step_grads(p, u, pos, neg, mt, g["margin"])
    _close(r["loss"], g["train_loss"])
    _close(r["scores"], g["train_scores"])
    _close(r["u_f"], g["train_u_f"])
    _close(r["pos_f"], g["train_pos_f"])
    _close(r["neg_f"], g["train_neg_f"])
    for k, v in g["grads"].items():
        _close(r["grads"][k], v, 5e-5)
    for k, v in g["grad_norms"].items():
        assert abs(r["grads"][k].double().norm() - v) <= 5e-5 * v + 1e-12, k
    for k, v in g["grad_samples"].items():
        gk = r["grads"][k].flatten()
        _close(gk[:: max(1, gk.numel() // 64)], v, 5e-5)
    for k, v in g["buffers_after"].items():
        if k in r["new_stats"]:
            if v.is_floating_point():
                _close(r["new_stats"][k], v)
            else:
                assert int(r["new_stats"][k]) == int(v)


@pytest.mark.parametrize("mt", O.MODEL_TYPES)
def test_oracle_matches_reference_eval(mt):
    g = _load("ref_%s.pt" % mt)
    p = fixtures.make_params(mt, seed=0, user_count=g["U"])
    u, pos, neg = fixtures.make_inputs(g["B"], g["N"], g["U"], seed=1)
    u[1] = u[0]
    with torch.no_grad():
        s, u_f, pos_f, neg_f = O.dcue_forward(p, u, pos, neg, mt, training=False)
        _close(s, g["eval_scores"])
        _close(u_f, g["eval_u_f"])
        _close(pos_f, g["eval_pos_f"])
        _close(neg_f, g["eval_neg_f"])
        _close(O.hinge_loss(s, g["margin"]), g["eval_loss"])
        item_f = O.tower_forward(p, pos, mt, training=False)
        _close(item_f, g["eval_item_f"])
        _close(O.user_forward(p, u), g["eval_user_f"])
        _close(O.cosine(g["eval_user_f"], g["eval_item_f"]), g["eval_sim"])


def test_hinge_known_answer():
    g = _load("ref_hinge_kat.pt")
    s = g["scores"].clone().requires_grad_(True)
    l = O.hinge_loss(s, 0.2)
    l.backward()
    assert abs(l.item() - 0.7) < 1e-6 and torch.allclose(l.detach(), g["loss"])
    assert torch.equal(s.grad, g["grad"])
    t = g["tie_scores"].clone().requires_grad_(True)
    O.hinge_loss(t, 0.2).backward()
    assert torch.equal(t.grad, g["tie_grad"])  # -0.5 at the exact tie


def test_score_hinge_fwdbwd_consistent():
    torch.manual_seed(0)
    B, N, F = 5, 4, 100
    u_f, feats = torch.randn(B, F), torch.randn(B * (1 + N), F)
    s, l, du, df = O.score_hinge_fwdbwd(u_f, feats, B, N, 0.2)
    s2, *_ = O.dcue_forward.__globals__["cosine"](u_f, feats[:B]), None
    assert s.shape == (B, N) and du.shape == u_f.shape and df.shape == feats.shape
    assert torch.isfinite(l)


def test_topk_oracle_matches_pairwise_sim():
    torch.manual_seed(0)
    uf, itf = torch.randn(7, 100), torch.randn(300, 100)
    v, i = O.topk_scores(uf, itf, 10)
    for r in range(7):
        sims = O.cosine(uf[r:r + 1].expand(300, -1), itf)
        vv, ii = torch.topk(sims, 10)
        assert torch.equal(ii, i[r])
        assert torch.allclose(vv, v[r], atol=1e-6)


def test_embedding_dense_grad():
    idx = torch.tensor([3, 1, 3, 0])
    g = torch.arange(8.0).view(4, 2)
    d = O.embedding_dense_grad(idx, g, 5)
    assert torch.equal(d[3], g[0] + g[2]) and torch.equal(d[2], torch.zeros(2))


@pytest.mark.parametrize("mt", ["truedcuemel1dbn", "truedcuemel1dresbn"])
def test_bn0_fold_is_exact_algebra(mt):
    """The B200 path feeds layer1 the normalised input and folds bn0's affine into layer1
    (weights * gamma, border-aware bias from beta; dW/dgamma/dbeta from G = sum dY*xhat).
    With identity roundings the folded oracle must reproduce the plain reference math."""
    p = fixtures.make_params(mt, seed=0, user_count=30)
    u, pos, neg = fixtures.make_inputs(4, 2, 30, seed=1)
    a = O.train_step_grads(p, u, pos, neg, mt, 0.2, dtype=torch.float64)
    b = O.train_step_grads(p, u, pos, neg, mt, 0.2, operand_dtype=torch.float64, grad_dtype=torch.float64,
                           dtype=torch.float64)
    assert abs(a["loss"].item() - b["loss"].item()) < 1e-12
    for k in a["grads"]:
        _close(b["grads"][k], a["grads"][k], 1e-9)
    for k in a["new_stats"]:
        if a["new_stats"][k].is_floating_point():
            _close(b["new_stats"][k], a["new_stats"][k], 1e-12)
