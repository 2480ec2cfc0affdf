"""Host-side logic without a GPU: every C-ABI call made by a full DCUE forward/backward is
checked against the prototypes parsed from include/dcue_b200.h (argument count and kinds).
No kernel runs; buffers are CPU tensors and outputs are garbage by construction."""
import ctypes
import importlib

import pytest
import torch

pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
L, ops = pkg._lib, pkg.ops


class _Recorder:
    def __init__(self):
        self.calls = []
        self.protos = L.parse_header()

    def call(self, name, *args):
        res, argtypes, names = self.protos[name]
        assert len(args) == len(argtypes), "%s: %d args, header has %d" % (name, len(args), len(argtypes))
        for a, t, n in zip(args, argtypes, names):
            if t is ctypes.c_void_p:
                assert a is None or (isinstance(a, int) and a >= 0), (name, n, a)
            elif t in (ctypes.c_float, ctypes.c_double):
                assert isinstance(a, (int, float)) and not isinstance(a, bool), (name, n, a)
            else:
                assert isinstance(a, int) and not isinstance(a, bool), (name, n, type(a))
        self.calls.append(name)

    def query(self, name, *args):
        res, argtypes, _ = self.protos[name]
        assert len(args) == len(argtypes), name
        return 1 << 16


@pytest.fixture
def dry(monkeypatch):
    rec = _Recorder()
    monkeypatch.setattr(L, "call", rec.call)
    monkeypatch.setattr(L, "query", rec.query)
    monkeypatch.setattr(L, "stream", lambda: 0)
    monkeypatch.setattr(ops, "_check_input", lambda x, name: x.contiguous())
    ops.clear_workspaces()
    yield rec
    ops.clear_workspaces()


@pytest.mark.parametrize("mt", ["truedcuemel1d", "truedcuemel1dres", "truedcuemel1dbn", "truedcuemel1dresbn"])
def test_forward_backward_marshalling(dry, monkeypatch, mt):
    monkeypatch.setattr(ops.UserTowerFn, "forward", staticmethod(_user_fwd_nocuda(ops.UserTowerFn.forward)))
    net = pkg.DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": 30, "model_type": mt})
    B, N = 3, 2
    u = torch.randint(0, 30, (B,))
    pos, neg = torch.randn(B, 128, 131), torch.randn(B, N, 128, 131)
    scores, u_f, pos_f, neg_f = net(u, pos, neg)
    assert scores.shape == (B, N) and u_f.shape == (B, 100) and pos_f.shape == (B, 100) and neg_f.shape == (B, N, 100)
    torch.nan_to_num(scores).sum().backward()
    for n, p in net.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape, n
    assert ("dcue_conv_pool_fwd" in dry.calls or "dcue_conv_pool_fwd_parts" in dry.calls) and "dcue_conv_wgrad" in dry.calls
    assert "dcue_scatter_add_rows" in dry.calls
    if mt.endswith("bn"):      # training-mode BatchNorm statistics take the fused finalisers
        assert "dcue_bn_stats_finalize" in dry.calls and "dcue_bn_bwd_finalize" in dry.calls
    # fused loss path
    net.zero_grad()
    loss = net.hinge_loss_step(u, pos, neg, margin=0.2)
    torch.nan_to_num(loss).backward()
    assert "dcue_score_hinge_fwdbwd" in dry.calls
    # eval mode, no grad: workspace returns to the pool
    net.eval()
    with torch.no_grad():
        out = net.conv(pos)
    assert out.shape == (B, 100)
    assert sum(len(v) for v in ops._POOL.values()) >= 1


def _user_fwd_nocuda(orig):
    def fwd(ctx, idx, table, *rest):  # rest = w1, b1, w2, b2, err
        class _T:  # pretend the table is on a CUDA device for the is_cuda guard only
            pass
        real_is_cuda = torch.Tensor.is_cuda
        try:
            torch.Tensor.is_cuda = property(lambda self: True)
            return orig(ctx, idx, table, *rest)
        finally:
            torch.Tensor.is_cuda = real_is_cuda
    return fwd


def test_geometry_rules():
    geo = ops.tower_geometry(131)
    assert [g["Lp"] for g in geo] == [136, 36, 12, 4] and [g["P"] for g in geo] == [33, 8, 2, 1]
    for g in geo:
        assert g["Lp"] % g["pool"] == 0 and g["Lp"] >= g["Lin"] + g["pad"]
        assert g["Lp"] - g["P"] * g["pool"] >= g["k"] - 1          # zero tail seen by taps / dgrad
    with pytest.raises(ValueError):
        ops.tower_geometry(40)


def test_unknown_model_type():
    with pytest.raises(ValueError):
        pkg.DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": 3, "model_type": "nope"})


def test_cpu_input_is_refused():
    net = pkg.DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": 3,
                       "model_type": "truedcuemel1dbn"})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(2, dtype=torch.int64), torch.randn(2, 128, 131), torch.randn(2, 1, 128, 131))
