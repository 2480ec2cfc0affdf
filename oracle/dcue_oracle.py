"""CPU oracle for the DCUE hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (plain torch CPU ops on explicit parameter
dicts, no nn.Module state) of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package never does.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, run in the build
container by ``oracle/make_golden.py`` and committed under ``tests/golden/``
(``tests/test_oracle_golden.py`` replays them on every CPU test run).

Reference anchors (all under /root/reference):
  * DCUENet.forward                dcrecommend/dcue/dcue.py:70-108
  * UserEmbeddings.forward         dcrecommend/dcue/embeddings/userembedding.py:33-44
  * TrueDcueNetMel1DBn.forward     dcrecommend/dcue/audiomodels/truedcuemel1dbn.py:77-101
  * TrueDcueNetMel1D.forward       dcrecommend/dcue/audiomodels/truedcuemel1d.py:69-87
  * TrueDcueNetMel1DRes.forward    dcrecommend/dcue/audiomodels/truedcuemel1dres.py:74-97
  * TrueDcueNetMel1DResBn.forward  dcrecommend/dcue/audiomodels/truedcuemel1dresbn.py:80-109
  * DCUE._loss_func                dcrecommend/nn/dcue.py:167-170
  * DCUE.predict (model.sim)       dcrecommend/nn/dcue.py:495-513

``operand_dtype`` / ``grad_dtype`` reproduce the roundings of the B200 path (and switch the k = 1 conv and the fc to
TF32-rounded operands, the single-pass tensor-core form those two GEMMs take there):
conv operands (activations entering layer1..4 and their weights) are rounded to
``operand_dtype`` (fp16 on B200) and the gradient entering each conv backward is
rounded to ``grad_dtype`` ("fp16_scaled": fp16 under a per-layer power-of-two scale);
accumulation stays in the working dtype.  With
both ``None`` the oracle is the reference's exact fp32 (or fp64) arithmetic.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

MODEL_TYPES = ("truedcuemel1d", "truedcuemel1dres", "truedcuemel1dbn", "truedcuemel1dresbn")
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
COS_EPS = 1e-8


def _round(x, dt):
    """Round to a 16-bit format.  dt == "fp16_scaled" is the B200 conv-backward operand: fp16 after
    a power-of-two scale chosen per layer so that no value overflows (dcue_grad_scale), i.e. an
    11-bit significand with effectively unbounded exponent."""
    if dt is None:
        return x
    if isinstance(dt, str):
        assert dt == "fp16_scaled"
        m, e = torch.frexp(x.double())
        return torch.ldexp(torch.round(m * 2048.0) / 2048.0, e).to(x.dtype)
    return x.to(dt).to(x.dtype)


def _round_tf32(x):
    """Round to TF32 (10 explicit mantissa bits) like PTX cvt.rna.tf32.f32: nearest, ties away from zero.  The B200 path runs
    the tower's k = 1 conv and fc as single-pass TF32 tensor-core GEMMs (csrc/linear.cu dcue_linear_*_tf32)."""
    f = x.detach().to(torch.float32).contiguous()
    bits = f.view(torch.int32)
    r = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    r = torch.where(torch.isfinite(f), r, f)
    return r.to(x.dtype)


class _RoundedLinear(torch.autograd.Function):
    """y = x @ w.T + b over the last dim with TF32-rounded operands; the gradient operand is rounded the same way in both
    backward GEMMs (dX = g_r @ w_r, dW = g_r.T @ x_r); the bias gradient uses the unrounded gradient."""

    @staticmethod
    def forward(ctx, x, w, b):
        xr, wr = _round_tf32(x), _round_tf32(w)
        ctx.save_for_backward(xr, wr)
        return F.linear(xr, wr, b)

    @staticmethod
    def backward(ctx, gy):
        xr, wr = ctx.saved_tensors
        gr = _round_tf32(gy)
        gx = gr @ wr
        gw = gr.reshape(-1, gr.shape[-1]).t() @ xr.reshape(-1, xr.shape[-1])
        gb = gy.reshape(-1, gy.shape[-1]).sum(0)
        return gx, gw, gb


class _RoundedConv(torch.autograd.Function):
    """conv1d whose operands are rounded like the B200 path (see module docstring)."""

    @staticmethod
    def forward(ctx, x, w, b, padding, op_dt, g_dt):
        xr, wr = _round(x, op_dt), _round(w, op_dt)
        ctx.save_for_backward(xr, wr)
        ctx.padding, ctx.g_dt = padding, g_dt
        return F.conv1d(xr, wr, b, padding=padding)

    @staticmethod
    def backward(ctx, gy):
        xr, wr = ctx.saved_tensors
        gr = _round(gy, ctx.g_dt)
        gx = torch.nn.grad.conv1d_input(xr.shape, wr, gr, padding=ctx.padding)
        gw = torch.nn.grad.conv1d_weight(xr, wr.shape, gr, padding=ctx.padding)
        gb = gy.sum(dim=(0, 2))  # bias grad is taken before the bf16 rounding
        return gx, gw, gb, None, None, None


class _FoldedBNConv(torch.autograd.Function):
    """layer1(bn0(x)) as the B200 path evaluates it.  The conv operand is u = x - running_mean (rounded to
    the operand dtype); the exact normalisation x-hat = r*u + shift, bn0.weight and bn0.bias are folded into
    the weights (W*gamma*r, rounded) and a border-aware bias; backward gets dW, dgamma, dbeta from
    G = sum dY*u (no data gradient).  Same function as the reference, different roundings."""

    @staticmethod
    def forward(ctx, u, w, b, gamma, beta, r, shift, padding, op_dt, g_dt):
        ur = _round(u, op_dt)
        wg = _round(w * (gamma * r)[None, :, None], op_dt)
        S, C, L = u.shape
        const_img = (beta + gamma * shift).view(1, C, 1).expand(1, C, L)
        y = F.conv1d(ur, wg, None, padding=padding) + F.conv1d(const_img, w, None, padding=padding) + b.view(1, -1, 1)
        ctx.save_for_backward(ur, w, gamma, beta, r, shift)
        ctx.padding, ctx.g_dt = padding, g_dt
        return y

    @staticmethod
    def backward(ctx, gy):
        ur, w, gamma, beta, r, shift = ctx.saved_tensors
        pad, k, L = ctx.padding, w.shape[2], ur.shape[2]
        gr = _round(gy, ctx.g_dt)
        Gu = torch.nn.grad.conv1d_weight(ur, w.shape, gr, padding=pad)
        tall = gy.sum(dim=(0, 2))
        T = tall[:, None].repeat(1, k)
        for j in range(k):
            for t in range(gy.shape[2]):
                if not 0 <= t + j - pad < L:
                    T[:, j] = T[:, j] - gr[:, :, t].sum(0)
        G = r[None, :, None] * Gu + shift[None, :, None] * T[:, None, :]      # sum dY * x-hat
        dW = gamma[None, :, None] * G + beta[None, :, None] * T[:, None, :]
        dgamma = (w * G).sum(dim=(0, 2))
        dbeta = (w * T[:, None, :]).sum(dim=(0, 2))
        return None, dW, tall, dgamma, dbeta, None, None, None, None, None


def _conv(x, w, b, padding, op_dt, g_dt):
    if op_dt is None and g_dt is None:
        return F.conv1d(x, w, b, padding=padding)
    return _RoundedConv.apply(x, w, b, padding, op_dt, g_dt)


def _bn(x, p, name, training, new_stats, affine=True):
    """BatchNorm1d over [S, C, L] (truedcuemel1dbn.py:24,30,...): batch statistics with
    biased variance when training, running statistics otherwise; running_var is updated
    with the unbiased variance, momentum 0.1, eps 1e-5."""
    w, b = p[name + ".weight"], p[name + ".bias"]
    rm, rv = p[name + ".running_mean"], p[name + ".running_var"]
    if training:
        n = x.shape[0] * x.shape[2]
        mean = x.mean(dim=(0, 2))
        var = x.var(dim=(0, 2), unbiased=False)
        if new_stats is not None:
            with torch.no_grad():
                new_stats[name + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.to(rm.dtype)
                new_stats[name + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * (var * n / max(n - 1, 1)).to(rv.dtype)
                new_stats[name + ".num_batches_tracked"] = p[name + ".num_batches_tracked"] + 1
    else:
        mean, var = rm.to(x.dtype), rv.to(x.dtype)
    xhat = (x - mean[None, :, None]) * torch.rsqrt(var[None, :, None] + BN_EPS)
    if not affine:
        return xhat
    return xhat * w[None, :, None] + b[None, :, None]


def tower_forward(p, x, model_type, training=True, prefix="conv.", operand_dtype=None,
                  grad_dtype=None, new_stats=None):
    """Song tower on x [S, 128, L] -> [S, F]  (reference files listed in the module docstring).

    Order per stage is conv -> maxpool -> relu -> (bn), exactly as the reference executes it.
    """
    if model_type not in MODEL_TYPES:
        raise ValueError("{} is not a recognized model type!".format(model_type))
    bn = model_type.endswith("bn")
    res = "res" in model_type
    q = {k[len(prefix):]: v for k, v in p.items() if k.startswith(prefix)}
    stats = None
    if new_stats is not None:
        stats = {}
    fold = bn and (operand_dtype is not None or grad_dtype is not None)
    if bn and not fold:
        x = _bn(x, q, "bn0", training, stats)
    elif bn:
        # B200 evaluation order: operand u = x - running_mean, statistics of the batch folded into layer1
        center = q["bn0.running_mean"].to(x.dtype)
        _bn(x, q, "bn0", training, stats, affine=False)          # running-statistics update only
        if training:
            mean_x, var = x.mean(dim=(0, 2)).detach(), x.var(dim=(0, 2), unbiased=False).detach()
        else:
            mean_x, var = center, q["bn0.running_var"].to(x.dtype)
        r0 = torch.rsqrt(var + BN_EPS)
        shift0 = -(mean_x - center) * r0
        x = x - center[None, :, None]
    tps = []
    for i, (pad, pool) in enumerate(((2, 4), (2, 4), (2, 4), (1, 2)), start=1):
        if fold and i == 1:
            x = _FoldedBNConv.apply(x, q["layer1.weight"], q["layer1.bias"], q["bn0.weight"], q["bn0.bias"], r0, shift0,
                                    pad, operand_dtype, grad_dtype)
        else:
            x = _conv(x, q["layer%d.weight" % i], q["layer%d.bias" % i], pad, operand_dtype, grad_dtype)
        x = F.max_pool1d(x, pool)
        x = F.relu(x)
        if bn:
            x = _bn(x, q, "bn%d" % i, training, stats)
        if res:
            tps.append(x.mean(dim=2, keepdim=True))  # AvgPool1d over the whole time extent
    tf32 = operand_dtype in (torch.float16, torch.bfloat16)          # B200 evaluation: single-pass TF32 for layer5 / fc
    if tf32:   # k = 1 conv == linear over the channel dim
        x = _RoundedLinear.apply(x.permute(0, 2, 1), q["layer5.weight"][:, :, 0], q["layer5.bias"]).permute(0, 2, 1)
    else:
        x = F.conv1d(x, q["layer5.weight"], q["layer5.bias"])
    x = F.relu(x)
    if bn:
        x = _bn(x, q, "bn5", training, stats)
    if res:
        x = torch.cat(tps + [x], dim=1)
    if stats is not None:
        for k, v in stats.items():
            new_stats[prefix + k] = v
    if tf32:
        out = _RoundedLinear.apply(x.permute(0, 2, 1), q["fc.weight"], q["fc.bias"])
    else:
        out = F.linear(x.permute(0, 2, 1), q["fc.weight"], q["fc.bias"])
    return out.reshape(out.shape[0], out.shape[-1]) if out.shape[1] == 1 else out


def user_forward(p, u, prefix="user_embd."):
    """UserEmbeddings.forward (userembedding.py:40-44): T[u] -> relu -> linear1 -> relu -> linear2."""
    h = p[prefix + "embeddings.weight"][u]
    h = F.relu(h)
    h = F.linear(h, p[prefix + "linear1.weight"], p[prefix + "linear1.bias"])
    h = F.relu(h)
    return F.linear(h, p[prefix + "linear2.weight"], p[prefix + "linear2.bias"])


def cosine(x, y, dim=1, eps=COS_EPS):
    """nn.CosineSimilarity(dim=1) as torch>=2.0 evaluates it (dcue.py:68): each vector is
    divided by max(||.||, eps) before the dot product."""
    xn = x / x.norm(dim=dim, keepdim=True).clamp_min(eps)
    yn = y / y.norm(dim=dim, keepdim=True).clamp_min(eps)
    return (xn * yn).sum(dim=dim)


def dcue_forward(p, u, pos, neg, model_type, training=True, operand_dtype=None, grad_dtype=None,
                 new_stats=None):
    """DCUENet.forward (dcue.py:70-108) -> (scores[B,N], u_f[B,F], pos_f[B,F], neg_f[B,N,F])."""
    u_f = user_forward(p, u)
    B, N = neg.shape[0], neg.shape[1]
    posneg = torch.cat([pos, neg.reshape(B * N, neg.shape[2], neg.shape[3])], dim=0)
    feats = tower_forward(p, posneg, model_type, training, "conv.", operand_dtype, grad_dtype, new_stats)
    pos_f = feats[:B]
    neg_f = feats[B:].reshape(B, N, -1)
    pos_s = cosine(u_f, pos_f)
    neg_s = cosine(u_f.unsqueeze(2), neg_f.permute(0, 2, 1))
    return pos_s.view(B, 1) - neg_s, u_f, pos_f, neg_f


def hinge_loss(scores, margin):
    """DCUE._loss_func (nn/dcue.py:167-170)."""
    return torch.max(torch.zeros_like(scores), margin - scores).sum(dim=1).mean()


def train_step_grads(p, u, pos, neg, model_type, margin, operand_dtype=None, grad_dtype=None,
                     dtype=torch.float32):
    """One forward + loss + backward.  Returns dict(loss, scores, u_f, pos_f, neg_f,
    grads{name: tensor}, new_stats{bn buffers after the step})."""
    q = {}
    for k, v in p.items():
        if v.is_floating_point():
            q[k] = v.detach().to(dtype).requires_grad_(not ("running_" in k))
        else:
            q[k] = v.detach().clone()
    new_stats = {}
    scores, u_f, pos_f, neg_f = dcue_forward(q, u, pos.to(dtype), neg.to(dtype), model_type, True,
                                              operand_dtype, grad_dtype, new_stats)
    loss = hinge_loss(scores, margin)
    names = [k for k, v in q.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [q[k] for k in names], allow_unused=True)
    return dict(loss=loss.detach(), scores=scores.detach(), u_f=u_f.detach(), pos_f=pos_f.detach(),
                neg_f=neg_f.detach(), grads={k: g for k, g in zip(names, grads) if g is not None},
                new_stats=new_stats)


def score_hinge_fwdbwd(u_f, feats, B, N, margin):
    """Scores + hinge loss and their gradient w.r.t. the feature vectors
    (dcue.py:93-106 + nn/dcue.py:167-170).  feats = [B pos rows; B*N neg rows]."""
    u_f = u_f.detach().clone().requires_grad_(True)
    feats = feats.detach().clone().requires_grad_(True)
    pos_s = cosine(u_f, feats[:B])
    neg_s = cosine(u_f.unsqueeze(2), feats[B:].reshape(B, N, -1).permute(0, 2, 1))
    scores = pos_s.view(B, 1) - neg_s
    loss = hinge_loss(scores, margin)
    du, df = torch.autograd.grad(loss, [u_f, feats])
    return scores.detach(), loss.detach(), du, df


def embedding_dense_grad(idx, grad_rows, num_rows):
    """autograd of nn.Embedding(sparse=False) (userembedding.py:27): dense [U,E] sum of rows."""
    out = torch.zeros(num_rows, grad_rows.shape[1], dtype=grad_rows.dtype)
    out.index_add_(0, idx, grad_rows)
    return out


def topk_scores(user_factors, item_factors, k, chunk=4096):
    """All-pairs form of DCUE.predict's model.sim(u, i) (nn/dcue.py:513) + top-k per user."""
    un = user_factors / user_factors.norm(dim=1, keepdim=True).clamp_min(COS_EPS)
    inn = item_factors / item_factors.norm(dim=1, keepdim=True).clamp_min(COS_EPS)
    vals, idxs = [], []
    for s in range(0, un.shape[0], chunk):
        sc = un[s:s + chunk] @ inn.T
        v, i = torch.topk(sc, k, dim=1)
        vals.append(v)
        idxs.append(i)
    return torch.cat(vals), torch.cat(idxs)
