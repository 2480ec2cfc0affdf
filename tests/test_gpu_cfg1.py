"""Parity at the BASELINE cfg1 shape (batch 64, 20 negatives, 20 000 users, truedcuemel1dbn) against

  (a) the UNMODIFIED reference's fp32 CPU step (tests/golden/ref_cfg1_truedcuemel1dbn.pt, written by
      ``oracle/make_golden.py --cfg1``), and
  (b) the oracle evaluated with the kernels' own operand roundings (decisions identical -> tight bounds),

for both 16-bit operand formats (fp16 = default, bf16 = ``DCUE_OPERAND=bf16``).  The test PRINTS the measured error table
(and writes it to gpurun_out/cfg1_error_table_<fmt>.json) and asserts bounds a little above the measured values.

What meets north_star's 1e-3 and what cannot (numbers: profiles/r02_cfg1_error_table.md):
  * loss, BatchNorm running statistics, user features (fp32 path): within 1e-3 for both formats;
  * song features / scores: fp16 ~2e-3, bf16 ~2e-2 of the largest value -- the operand rounding itself (2^-11 / 2^-8 per
    element through four conv layers); fp16 carries exactly TF32's 11 significant bits;
  * tower gradients: 3-8 % (fp16), 10-25 % (bf16) per tensor in l2.  This is not kernel error: a score that moves by 1e-3
    flips ~0.3 % of the hinge / max-pool / ReLU decisions of the batch, a flipped decision changes its gradient
    contribution by 100 %, so the relative gradient error is ~sqrt(flip fraction) whatever the batch size.  The CPU oracle
    with the same roundings shows the same distance to fp32 (oracle/make_golden.py docstring), and against THAT oracle the
    kernels agree to the bounds of level (b).  Only fp32-exact scores (error < 1e-6) could give 1e-3 gradients.
"""
import importlib
import json
import os

import pytest
import torch

from oracle import dcue_oracle as O
from oracle import fixtures

pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(__file__))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_cfg1_truedcuemel1dbn.pt")

# (vs fp32 reference, vs same-rounding oracle) bounds per operand format
BOUNDS = {
    # measured on a B200 (profiles/r02_cfg1_error_table.md): fp16 loss 3.9e-5, features 1.95e-3, scores 1.0e-3, tower gradients
    # l2 <= 0.067 / cos >= 0.9977, user MLP + table 2.1e-3, buffers 2.7e-5; vs its own oracle 5.6e-6 / 4.3e-4 / 0.051 / 0.9987
    "f16": dict(loss=1e-3, feat=5e-3, score_abs=3e-3, grad_l2=0.15, grad_cos=0.99, mlp_l2=6e-3, table_l2=6e-3, buf=2e-4,
                o_loss=1e-4, o_feat=1e-3, o_grad_l2=0.10, o_grad_cos=0.997),
    # bf16: loss 4.1e-5, features 1.7e-2, scores 1.1e-2, tower gradients l2 <= 0.224 / cos >= 0.975, MLP + table 1.7e-2,
    # buffers 2.9e-4; vs its own oracle 1.3e-4 / 2.0e-3 / 0.012 / 0.99993
    "bf16": dict(loss=1e-3, feat=4e-2, score_abs=2.5e-2, grad_l2=0.40, grad_cos=0.95, mlp_l2=4e-2, table_l2=4e-2, buf=1e-3,
                 o_loss=5e-4, o_feat=6e-3, o_grad_l2=0.05, o_grad_cos=0.998),
}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300)).item()


@pytest.mark.parametrize("fmt", ["f16", "bf16"])
def test_cfg1_train_step_error_table(fmt, monkeypatch):
    monkeypatch.setenv("DCUE_OPERAND", fmt)
    g = torch.load(GOLD, weights_only=False)
    mt, B, N, U, margin = g["model_type"], g["B"], g["N"], g["U"], g["margin"]
    params = fixtures.make_params(mt, seed=0, user_count=U)
    u, pos, neg = fixtures.make_inputs(B, N, U, seed=1)
    u[1] = u[0]
    assert torch.equal(u, g["u"])                                    # gather indices: bit-exact inputs
    net = pkg.DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": U, "model_type": mt})
    net.load_state_dict(params)
    net = net.to(DEV).train()
    loss, scores, u_f, pos_f, neg_f = net.hinge_loss_step(u.to(DEV), pos.to(DEV), neg.to(DEV), margin, return_all=True)
    loss.backward()
    torch.cuda.synchronize()
    bd = BOUNDS[fmt]
    tab = {"fmt": fmt, "shape": "cfg1 B=%d N=%d U=%d S=%d" % (B, N, U, B * (1 + N))}

    # ---------------- (a) against the reference's fp32 step
    a = tab["vs_reference_fp32"] = {}
    a["loss_rel"] = abs(loss.item() - g["train_loss"].item()) / abs(g["train_loss"].item())
    a["u_f_relmax"] = rel(u_f, g["train_u_f"])
    a["pos_f_relmax"], a["neg_f_relmax"] = rel(pos_f, g["train_pos_f"]), rel(neg_f, g["train_neg_f"])
    a["scores_absmax"] = (scores.cpu() - g["train_scores"]).abs().max().item()
    hinge_ref = (g["train_scores"] < margin)
    a["hinge_decisions_flipped"] = int(((scores.cpu() < margin) != hinge_ref).sum())
    a["hinge_decisions_total"] = int(hinge_ref.numel())
    a["grads"] = {}
    for k, v in g["grads"].items():
        got = net.get_parameter(k).grad
        a["grads"][k] = {"l2": l2(got, v), "relmax": rel(got, v), "cos": cos(got, v)}
    tg = net.user_embd.embeddings.weight.grad
    a["table_rows_l2"] = l2(tg[g["table_grad_rows"].to(DEV)], g["table_grad"])
    a["table_norm_rel"] = abs(tg.double().norm().item() - g["table_grad_norm"].item()) / g["table_grad_norm"].item()
    untouched = torch.ones(U, dtype=torch.bool, device=DEV)
    untouched[g["table_grad_rows"].to(DEV)] = False
    a["table_untouched_rows_nonzero"] = int((tg[untouched] != 0).sum())
    a["buffers"] = {}
    bufs = dict(net.named_buffers())
    for k, v in g["buffers_after"].items():
        if v.is_floating_point():
            a["buffers"][k] = rel(bufs[k], v)
        else:
            assert int(bufs[k]) == int(v), k

    # ---------------- (b) against the oracle with the kernels' roundings (fp64 accumulation)
    od = torch.float16 if fmt == "f16" else torch.bfloat16
    gd = "fp16_scaled" if fmt == "f16" else torch.bfloat16     # the gradient operand takes the forward operand's format
    ref = O.train_step_grads(params, u, pos, neg, mt, margin, operand_dtype=od, grad_dtype=gd, dtype=torch.float64)
    b = tab["vs_rounded_oracle"] = {}
    b["loss_rel"] = abs(loss.item() - ref["loss"].item()) / abs(ref["loss"].item())
    b["pos_f_relmax"], b["neg_f_relmax"] = rel(pos_f, ref["pos_f"]), rel(neg_f, ref["neg_f"])
    b["hinge_decisions_flipped"] = int(((scores.cpu() < margin) != (ref["scores"] < margin)).sum())
    b["grads"] = {}
    for k, v in ref["grads"].items():
        if k == "user_embd.embeddings.weight":
            continue
        got = net.get_parameter(k).grad
        b["grads"][k] = {"l2": l2(got, v), "cos": cos(got, v)}

    print("\n" + json.dumps(tab, indent=1))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "cfg1_error_table_%s.json" % fmt), "w") as f:
        json.dump(tab, f, indent=1)

    assert a["loss_rel"] < bd["loss"]
    assert a["u_f_relmax"] < 1e-5
    assert a["pos_f_relmax"] < bd["feat"] and a["neg_f_relmax"] < bd["feat"]
    assert a["scores_absmax"] < bd["score_abs"]
    # a hinge decision that flips (a score within the operand rounding of the margin; 0-2 of 1 280 here, depending on the
    # accumulation order of the day) switches one (user, negative) pair's whole gradient on or off: ~1e-2 of the user-side
    # gradients per flip at this batch size -- allowed for explicitly, and reported in the table
    flips = a["hinge_decisions_flipped"]
    for k, e in a["grads"].items():
        if k.startswith("user_embd."):
            assert e["l2"] < bd["mlp_l2"] + 1e-2 * flips, (k, e, flips)
        else:
            assert e["l2"] < bd["grad_l2"] and e["cos"] > bd["grad_cos"], (k, e)
    assert a["table_rows_l2"] < bd["table_l2"] + 1e-2 * flips and a["table_norm_rel"] < bd["table_l2"]
    assert a["table_untouched_rows_nonzero"] == 0
    for k, e in a["buffers"].items():
        assert e < bd["buf"], (k, e)
    assert b["loss_rel"] < bd["o_loss"]
    assert b["pos_f_relmax"] < bd["o_feat"] and b["neg_f_relmax"] < bd["o_feat"]
    for k, e in b["grads"].items():
        assert e["l2"] < bd["o_grad_l2"] and e["cos"] > bd["o_grad_cos"], (k, e)


def test_cfg1_eval_forward_vs_reference():
    g = torch.load(GOLD, weights_only=False)
    mt, B, N, U = g["model_type"], g["B"], g["N"], g["U"]
    params = fixtures.make_params(mt, seed=0, user_count=U)
    u, pos, neg = fixtures.make_inputs(B, N, U, seed=1)
    u[1] = u[0]
    net = pkg.DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": U, "model_type": mt})
    net.load_state_dict(params)
    net = net.to(DEV).eval()
    with torch.no_grad():
        s, uf, pf, nf = net(u.to(DEV), pos.to(DEV), neg.to(DEV))
    assert rel(pf, g["eval_pos_f"]) < 5e-3 and rel(nf, g["eval_neg_f"]) < 5e-3
    assert (s.cpu() - g["eval_scores"]).abs().max() < 3e-3
    loss = torch.max(torch.zeros_like(s), g["margin"] - s).sum(dim=1).mean().item()
    assert abs(loss - g["eval_loss"].item()) < 1e-3 * abs(g["eval_loss"].item())
