"""Generate tests/golden/*.pt by running the UNMODIFIED reference (imported read-only from
/root/reference) on the seeded fixtures of oracle/fixtures.py.

Run in the build container only (the GPU box has no /root/reference):
    python oracle/make_golden.py
The committed .pt files pin oracle/dcue_oracle.py (tests/test_oracle_golden.py).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

from oracle import fixtures  # noqa: E402

from dcrecommend.dcue.dcue import DCUENet  # noqa: E402  (reference)
from dcrecommend.nn.dcue import DCUE  # noqa: E402  (reference trainer, for _loss_func)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
B, N, U, MARGIN = 6, 3, 50, 0.2


def run(model_type):
    p = fixtures.make_params(model_type, seed=0, user_count=U)
    u, pos, neg = fixtures.make_inputs(B, N, U, seed=1)
    u[1] = u[0]  # a duplicate row for the dense table gradient
    m = DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": U,
                 "model_type": model_type})
    m.load_state_dict(p)
    trainer = DCUE(margin=MARGIN)
    out = {"model_type": model_type, "B": B, "N": N, "U": U, "margin": MARGIN, "u": u}
    # train-mode step (reference nn/dcue.py:202-208)
    m.train()
    m.zero_grad()
    scores, u_f, pos_f, neg_f = m(u, pos, neg)
    loss = trainer._loss_func(scores)
    loss.backward()
    out.update(train_loss=loss.detach(), train_scores=scores.detach(), train_u_f=u_f.detach(),
               train_pos_f=pos_f.detach(), train_neg_f=neg_f.detach())
    out["grads"] = {k: v.grad.detach().clone() for k, v in m.named_parameters()
                    if v.numel() <= 4096 or k == "user_embd.embeddings.weight"}
    out["grad_norms"] = {k: v.grad.detach().double().norm() for k, v in m.named_parameters()}
    out["grad_samples"] = {k: v.grad.detach().flatten()[:: max(1, v.numel() // 64)].clone()
                           for k, v in m.named_parameters()}
    out["buffers_after"] = {k: v.detach().clone() for k, v in m.named_buffers()}
    # eval-mode forward (reference nn/dcue.py:220-262)
    m.load_state_dict(p)
    m.eval()
    with torch.no_grad():
        scores, u_f, pos_f, neg_f = m(u, pos, neg)
        out.update(eval_scores=scores, eval_u_f=u_f, eval_pos_f=pos_f, eval_neg_f=neg_f,
                   eval_loss=trainer._loss_func(scores))
        out["eval_item_f"] = m.conv(pos)          # _item_factors path (nn/dcue.py:663)
        out["eval_user_f"] = m.user_embd(u)       # _user_factors path (nn/dcue.py:638)
        out["eval_sim"] = m.sim(out["eval_user_f"], out["eval_item_f"])  # predict (nn/dcue.py:513)
    return out


def run_cfg1(model_type="truedcuemel1dbn", B1=64, N1=20, U1=20000):
    """BASELINE configs[0] (cfg1) shape: batch 64, 20 negatives, 20 000 users, fp32 CPU reference.  Stores the loss, scores,
    feature vectors, EVERY tower / user-MLP gradient in full, the touched rows of the dense table gradient and the
    BatchNorm buffers after the step."""
    p = fixtures.make_params(model_type, seed=0, user_count=U1)
    u, pos, neg = fixtures.make_inputs(B1, N1, U1, seed=1)
    u[1] = u[0]
    m = DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": U1, "model_type": model_type})
    m.load_state_dict(p)
    trainer = DCUE(margin=MARGIN)
    m.train()
    m.zero_grad()
    scores, u_f, pos_f, neg_f = m(u, pos, neg)
    loss = trainer._loss_func(scores)
    loss.backward()
    out = {"model_type": model_type, "B": B1, "N": N1, "U": U1, "margin": MARGIN, "u": u,
           "train_loss": loss.detach(), "train_scores": scores.detach(), "train_u_f": u_f.detach(),
           "train_pos_f": pos_f.detach(), "train_neg_f": neg_f.detach()}
    out["grads"] = {k: v.grad.detach().clone() for k, v in m.named_parameters() if k != "user_embd.embeddings.weight"}
    tg = m.user_embd.embeddings.weight.grad
    rows = torch.unique(u)
    out["table_grad_rows"] = rows
    out["table_grad"] = tg[rows].clone()
    out["table_grad_norm"] = tg.double().norm()
    out["buffers_after"] = {k: v.detach().clone() for k, v in m.named_buffers()}
    m.load_state_dict(p)
    m.eval()
    with torch.no_grad():
        scores, u_f, pos_f, neg_f = m(u, pos, neg)
        out.update(eval_scores=scores, eval_pos_f=pos_f, eval_neg_f=neg_f, eval_loss=trainer._loss_func(scores))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    if "--cfg1" in sys.argv:
        torch.save(run_cfg1(), os.path.join(OUT, "ref_cfg1_truedcuemel1dbn.pt"))
        print("wrote cfg1 golden")
        return
    torch.set_num_threads(1)  # deterministic summation order
    for mt in ("truedcuemel1d", "truedcuemel1dres", "truedcuemel1dbn", "truedcuemel1dresbn"):
        torch.save(run(mt), os.path.join(OUT, "ref_%s.pt" % mt))
        print("wrote", mt)
    # known-answer test implied by the reference's scratch block (dcue/dcue.py:152-161):
    # scores [[-1,2,2],[0,2,2]], margin 0.2 -> _loss_func = 0.7
    s = torch.tensor([[-1.0, 2.0, 2.0], [0.0, 2.0, 2.0]], requires_grad=True)
    l = DCUE(margin=0.2)._loss_func(s)
    l.backward()
    # tie at the margin: torch.max splits the gradient
    t = torch.tensor([[0.2, 0.1, 0.3]], requires_grad=True)
    lt = DCUE(margin=0.2)._loss_func(t)
    lt.backward()
    torch.save({"scores": s.detach(), "loss": l.detach(), "grad": s.grad, "tie_scores": t.detach(),
                "tie_loss": lt.detach(), "tie_grad": t.grad}, os.path.join(OUT, "ref_hinge_kat.pt"))
    print("wrote hinge KAT", float(l), t.grad)
    optim_goldens()


def optim_goldens():
    """Trajectories of the reference's Ranger and CyclicLRWithRestarts (host-side components the
    drop-in trainer re-implements) on a tiny deterministic problem."""
    import contextlib
    import io
    import warnings
    warnings.simplefilter("ignore")
    from dcrecommend.optim.cyclic_scheduler import CyclicLRWithRestarts
    from dcrecommend.optim.ranger import Ranger
    g = torch.Generator().manual_seed(5)
    w0, A, b = torch.randn(7, 5, generator=g), torch.randn(5, 5, generator=g), torch.randn(7, generator=g)
    w = torch.nn.Parameter(w0.clone())
    bb = torch.nn.Parameter(b.clone())
    with contextlib.redirect_stdout(io.StringIO()):
        opt = Ranger([w, bb], lr=1e-2, alpha=0.5, k=6, N_sma_threshhold=5, betas=(0.9, 0.99), eps=1e-5, weight_decay=1e-2)
    traj = []
    for _ in range(40):
        opt.zero_grad()
        loss = ((w @ A).tanh().sum(1) + bb).pow(2).sum()
        loss.backward()
        opt.step()
        traj.append(torch.cat([w.detach().flatten(), bb.detach()]).clone())
    p = torch.nn.Parameter(torch.zeros(3))
    sgd = torch.optim.SGD([p], lr=0.1, weight_decay=0.01)
    sch = CyclicLRWithRestarts(sgd, 4, 18, restart_period=2, t_mult=2, policy="cosine")
    lrs = []
    for epoch in range(9):
        sch.step()
        for _ in range(4):
            sch.batch_step()
            lrs.append((sgd.param_groups[0]["lr"], sgd.param_groups[0]["weight_decay"]))
    torch.save({"w0": w0, "A": A, "b": b, "ranger_traj": torch.stack(traj), "sched": torch.tensor(lrs, dtype=torch.float64)},
               os.path.join(OUT, "ref_optim.pt"))
    print("wrote optim goldens")


if __name__ == "__main__":
    main()
