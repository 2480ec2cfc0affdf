"""Seeded synthetic parameters / inputs shared by the golden generator, the tests, smoke()
and bench.py.  TEST INFRASTRUCTURE (lives with the oracle; the product never imports it).

Parameters are drawn key by key from one CPU generator, so the reference module
(``oracle/make_golden.py``), the oracle and the B200 module can all be loaded with
bit-identical values without depending on nn.Module init order.
Key names / shapes are the reference's ``state_dict``.
"""
from __future__ import annotations

import torch

N_MELS = 128
N_FRAMES = 131


def param_shapes(model_type, feature_dim=100, conv_hidden=128, user_embdim=300, user_count=50):
    F_, H, E, U = feature_dim, conv_hidden, user_embdim, user_count
    bn = model_type.endswith("bn")
    res = "res" in model_type
    shapes = {}

    def add_bn(name, c):
        shapes["conv.%s.weight" % name] = (c,)
        shapes["conv.%s.bias" % name] = (c,)
        shapes["conv.%s.running_mean" % name] = (c,)
        shapes["conv.%s.running_var" % name] = (c,)
        shapes["conv.%s.num_batches_tracked" % name] = ()

    if bn:
        add_bn("bn0", N_MELS)
    for i, (cin, cout, k) in enumerate(((N_MELS, H, 4), (H, H, 4), (H, H, 4), (H, H, 2), (H, F_, 1)), start=1):
        shapes["conv.layer%d.weight" % i] = (cout, cin, k)
        shapes["conv.layer%d.bias" % i] = (cout,)
        if bn:
            add_bn("bn%d" % i, cout)
    shapes["conv.fc.weight"] = (F_, 4 * H + F_ if res else F_)
    shapes["conv.fc.bias"] = (F_,)
    shapes["user_embd.embeddings.weight"] = (U, E)
    shapes["user_embd.linear1.weight"] = (E, E)
    shapes["user_embd.linear1.bias"] = (E,)
    shapes["user_embd.linear2.weight"] = (F_, E)
    shapes["user_embd.linear2.bias"] = (F_,)
    return shapes


def make_params(model_type, seed=0, **kw):
    g = torch.Generator().manual_seed(seed)
    p = {}
    for k, shp in param_shapes(model_type, **kw).items():
        if k.endswith("num_batches_tracked"):
            p[k] = torch.tensor(3, dtype=torch.int64)
        elif k.endswith("running_var"):
            p[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("running_mean"):
            p[k] = 0.2 * torch.randn(shp, generator=g)
        elif ".bn" in k and k.endswith("weight"):
            p[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            p[k] = 0.1 * torch.randn(shp, generator=g)
        elif k == "user_embd.embeddings.weight":
            p[k] = torch.randn(shp, generator=g)
        else:  # conv / linear weights: fan-in scaled
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            p[k] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
    return p


def make_inputs(B, N, user_count, seed=1, zipf=False, frames=N_FRAMES):
    """u int64 [B], pos f32 [B,128,L], neg f32 [B,N,128,L]  (SURVEY.md §8d)."""
    g = torch.Generator().manual_seed(seed)
    if zipf:  # duplicate-heavy indices to stress the segment scatter-add
        r = torch.rand(B, generator=g)
        u = (user_count * r ** 4).long().clamp_(0, user_count - 1)
    else:
        u = torch.randint(0, user_count, (B,), generator=g)
    pos = torch.randn(B, N_MELS, frames, generator=g)
    neg = torch.randn(B, N, N_MELS, frames, generator=g)
    return u, pos, neg
