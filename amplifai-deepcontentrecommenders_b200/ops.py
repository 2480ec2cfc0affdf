"""Host-side orchestration of the DCUE hot path: thin ``torch.autograd.Function``s that hand raw
device pointers of PyTorch-owned buffers to the C-ABI kernels (``_lib``).  No ATen compute and
no CPU fallback here: PyTorch provides memory, streams and autograd bookkeeping only.

Reference call sites replaced
  song tower   dcrecommend/dcue/audiomodels/truedcuemel1d{,bn,res,resbn}.py  forward + autograd
  user tower   dcrecommend/dcue/embeddings/userembedding.py:33-44            forward + autograd
  scoring      dcrecommend/dcue/dcue.py:93-106, dcrecommend/nn/dcue.py:167-170
"""
from __future__ import annotations

import os

import torch

from . import _lib as L

N_MELS = 128
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
COS_EPS = 1e-8

#  (k, pad, pool) of layer1..layer4 (truedcuemel1dbn.py:25-54)
STAGES = ((4, 2, 4), (4, 2, 4), (4, 2, 4), (2, 1, 2))


def _roundup(a, b):
    return (a + b - 1) // b * b


def conv_impl():
    """tcgen05 path unless DCUE_CONV_IMPL=simt (validator) is requested."""
    return L.IMPL_SIMT if os.environ.get("DCUE_CONV_IMPL", "tc").lower() == "simt" else L.IMPL_TC


def fused_wgrad():
    """Layer 1 backward as ONE tcgen05 kernel (BatchNorm-backward + unpool + weight gradient, no dY panel in HBM)
    unless DCUE_FUSED_WGRAD=0 selects the separate unpool -> panel -> wgrad kernels."""
    return os.environ.get("DCUE_FUSED_WGRAD", "1") != "0"


def fused_finalize():
    """BatchNorm statistic partials -> (peer all-reduce) -> finalize in ONE kernel per layer (DCUE_FUSED_FINALIZE=0: the
    separate reduce / all-reduce / finalize launches)."""
    return os.environ.get("DCUE_FUSED_FINALIZE", "1") != "0"


def _peer_args(dp):
    """(bufs, signals, counter, rank, world) of the NVLink peer all-reduce for the fused finalisers; world 1 = no exchange."""
    if dp is None or dp.world_size == 1:
        return (None, None, None, 0, 1)
    pr = dp._peer
    return (pr.hdl.buffer_ptrs_dev, pr.hdl.signal_pad_ptrs_dev, pr.counter.data_ptr(), pr.rank, pr.world)


def tower_linear(name):
    """The song tower's k = 1 conv and fc run as single-pass TF32 tensor-core GEMMs (their inputs already carry the 16-bit
    operand rounding of the conv layers); DCUE_TOWER_TF32=0 selects the fp32-accurate 3xTF32 form."""
    return name + "_tf32" if os.environ.get("DCUE_TOWER_TF32", "1") != "0" else name


def dgrad_stats():
    """BatchNorm-backward reductions of stage i-1 taken in the epilogue of stage i's data-gradient kernel
    (DCUE_DGRAD_STATS=1).  Off by default: measured slower than the separate dcue_bn_bwd_reduce sweep
    (layer-2 data gradient 124 -> ~600 us against 103 us saved)."""
    return os.environ.get("DCUE_DGRAD_STATS", "0") != "0"


def operand_fmt():
    """16-bit format of the forward conv operands: fp16 (default) or bf16 (DCUE_OPERAND=bf16)."""
    return L.FMT_BF16 if os.environ.get("DCUE_OPERAND", "f16").lower() == "bf16" else L.FMT_F16


def tower_geometry(frames):
    """Flat padded row geometry of the four conv stages for `frames` input frames."""
    geo, lin = [], frames
    for k, pad, pool in STAGES:
        lout = lin + 2 * pad - k + 1
        P = lout // pool
        if P < 1:
            raise ValueError("input of %d frames is too short for the DCUE tower" % frames)
        # rows per spectrogram: all data rows, and a zero tail of >= k-1 rows after the last live
        # conv output so neither the taps nor the dgrad ever see a neighbouring spectrogram
        lp = _roundup(max(lin + pad, P * pool + k - 1), pool)
        geo.append(dict(k=k, pad=pad, pool=pool, Lin=lin, Lout=lout, P=P, Lp=lp))
        lin = P
    if lin != 1:
        raise ValueError("the DCUE tower needs an input length that pools down to 1 frame (got %d -> %d)"
                         % (frames, lin))
    return geo


class Panel:
    """16-bit [rows, 128] activation matrix in the panel layout, zero-initialised (padding rows
    are never written afterwards)."""

    def __init__(self, S, Lp, device):
        self.rows_total = S * Lp
        self.panel_rows = L.FRONT_HALO + _roundup(max(self.rows_total, 1), 128) + L.BACK_HALO
        self.buf = torch.zeros(16 * self.panel_rows * 8, dtype=torch.int16, device=device)
        self.base = self.buf.data_ptr() + L.FRONT_HALO * 16


class TowerWorkspace:
    """All device buffers of one tower forward(+backward) for a given (S, frames)."""

    def __init__(self, S, frames, H, F, res, device):
        self.key = (S, frames, H, F, res, str(device))
        self.S, self.geo, self.device = S, tower_geometry(frames), device
        f32 = dict(dtype=torch.float32, device=device)
        self.X = [Panel(S, g["Lp"], device) for g in self.geo]
        self.z = [torch.empty(S * g["P"], H, **f32) for g in self.geo]
        self.code = [torch.empty(S * g["P"], H, dtype=torch.uint8, device=device) for g in self.geo]
        self.y4 = torch.empty(S, H, **f32)
        self.z5 = torch.empty(S, F, **f32)
        self.fc_in = torch.empty(S, 4 * H + F if res else F, **f32)
        self.sums = torch.zeros(6, 2 * 128, dtype=torch.float64, device=device)
        self.bnp = torch.zeros(6, 4, 128, **f32)        # scale, shift, mean, rstd per BN layer
        self.tapb = torch.zeros(4, 128, **f32)          # layer-1 per-tap constants of the folded bn0 shift
        self.ticket = torch.zeros(4, dtype=torch.int32, device=device)    # block ticket of the fused finalisers (self-resetting)
        self.zero128 = torch.zeros(128, **f32)
        self.one128 = torch.ones(128, **f32)
        self.wp = [torch.empty(128 * g["k"] * 128, dtype=torch.int16, device=device) for g in self.geo]
        nbytes = max(L.query("dcue_conv_ws_bytes", L.IMPL_TC, S, g["Lp"], 4, 128, 128) for g in self.geo)
        nbytes = max(nbytes, L.query("dcue_ncl_stats_ws_bytes", 128), L.query("dcue_bn_bwd_ws_bytes", 128),
                     L.query("dcue_conv_wgrad_unpool_ws_bytes", 4),
                     L.query("dcue_linear_wgrad_ws_bytes", S, 4 * H + F, F),
                     L.query("dcue_linear_wgrad_ws_bytes", S, H, F))
        self.scratch = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self._bwd = None

    def dy_panel(self, i):
        """Gradient operand panel of stage i (0-based); layer 1's is never needed on the fused wgrad path (749 MB at cfg2)."""
        b = self.bwd()
        if b["dY"][i] is None:
            b["dY"][i] = Panel(self.S, self.geo[i]["Lp"], self.device)
        return b["dY"][i]

    def bwd(self):
        """Backward-only buffers, created on first use."""
        if self._bwd is None:
            S, dev = self.S, self.device
            f32 = dict(dtype=torch.float32, device=dev)
            b = {}
            b["dY"] = [None] * len(self.geo)       # 16-bit gradient operand panels, created on first use (dy_panel)
            b["dx"] = [torch.empty(S * g["Lin"], 128, **f32) for g in self.geo]
            b["wpd"] = [torch.empty(128 * g["k"] * 128, dtype=torch.int16, device=dev) for g in self.geo]
            b["bsum"] = torch.zeros(2 * 128, dtype=torch.float64, device=dev)
            b["dsums"] = torch.zeros(6, 2 * 128, dtype=torch.float64, device=dev)
            b["E"] = torch.zeros(L.lib().dcue_panel_row_sums_parts(), 4, 128, **f32)   # border row sums, per slice
            b["amax"] = torch.zeros(6, **f32)
            b["gscale"] = torch.ones(6, 2, **f32)
            self._bwd = b
        return self._bwd


_POOL = {}


def _acquire(S, frames, H, F, res, device):
    key = (S, frames, H, F, res, str(device))
    free = _POOL.get(key)
    if free:
        return free.pop()
    return TowerWorkspace(S, frames, H, F, res, device)


def _release(ws):
    _POOL.setdefault(ws.key, []).append(ws)


def clear_workspaces():
    _POOL.clear()


def _check_input(x, name):
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the DCUE B200 path has no CPU fallback" % name)
    if x.dtype != torch.float32:
        raise TypeError("%s must be float32" % name)
    return x.contiguous()


# ---------------------------------------------------------------------------------------------------------------------------
# The user tower's backward (gather-side scatter + 4 small GEMMs: 66 us of latency-bound launches) overlaps the first kernels
# of the song tower's backward on a side stream.  It is forked only when DCUENet evaluated the user tower AFTER a song tower
# that needs gradients (autograd then runs the user tower's backward first and the song tower's backward -- which joins the
# stream at its end -- later in the same pass) and when autograd will ADOPT every gradient (all .grad None): an accumulating
# `grad += g` would run on the main stream without waiting for the side stream.  DCUE_USER_BWD_STREAM=0 disables it.
_BWD_SIDE = {"stream": None, "pending": None}


def _bwd_side_stream(device):
    st = _BWD_SIDE["stream"]
    if st is None or st.device != device:
        st = _BWD_SIDE["stream"] = torch.cuda.Stream(device=device)
    return st


def join_backward_side():
    """The current stream waits for a user-tower backward that is still running on the side stream (no-op otherwise).
    Called at the end of SongTowerFn.backward; also by GraphedTrainStep, DataParallelDCUE.reduce_gradients, the fused
    optimizers and the next forward, so that no consumer can get ahead of it."""
    pend = _BWD_SIDE["pending"]
    if pend is not None:
        _BWD_SIDE["pending"] = None
        # the consumer's stream on the side stream's device (the caller may have another device current)
        torch.cuda.current_stream(pend[0].device).wait_stream(pend[0])


class SongTowerFn(torch.autograd.Function):
    """feats[S,F] = tower(cat(pos, neg))  without materialising the concatenation."""

    @staticmethod
    def forward(ctx, pos, neg, src, mod, training, *params):
        """Dense feed: pos [B,128,L] (+ neg [.., 128, L]).  Index feed: pos = resident pool [n_songs,128,T],
        neg = None, src = (idx int64 [S], off int32 [S] or None, frames, err_flag int32 [1])."""
        join_backward_side()
        has_bn, res = mod._has_bn, mod._res
        H, F = mod.hidden_size, mod.output_size
        if H != 128:
            raise NotImplementedError("conv_hidden must be 128 for the B200 tower kernels (got %d)" % H)
        if F > 128 or F % 4:
            raise NotImplementedError("feature_dim must be a multiple of 4 and <= 128 (got %d)" % F)
        pos = _check_input(pos, "pos")
        S_pos, C, frames = pos.shape
        if C != N_MELS:
            raise ValueError("expected %d mel bins, got %d" % (N_MELS, C))
        if src is not None:
            idx, off, frames, err = src
            n_songs, _, T = pos.shape
            if idx.dtype != torch.int64 or (off is not None and off.dtype != torch.int32):
                raise TypeError("song indices must be int64 and crop offsets int32")
            idx = idx.to(pos.device).contiguous().view(-1)
            off = None if off is None else off.to(pos.device).contiguous().view(-1)
            S_pos, S_neg, neg = idx.numel(), 0, None
        else:
            if neg is not None:
                neg = _check_input(neg, "neg").view(-1, C, frames)
            S_neg = 0 if neg is None else neg.shape[0]
        S = S_pos + S_neg
        dev = pos.device
        st = L.stream()
        impl, fmt = conv_impl(), operand_fmt()
        ws = _acquire(S, frames, H, F, res, dev)
        geo = ws.geo
        scratch, nscr = ws.scratch.data_ptr(), ws.scratch.numel()
        P = dict(zip(mod._param_names, params))
        dp = mod._dp
        world = 1 if dp is None else dp.world_size

        # training-mode statistics take the fused finaliser (partials -> [peer all-reduce] -> scale/shift/mean/rstd + running
        # statistics in one launch) unless the data-parallel group has no peer memory (then: reduce, NCCL, finalize)
        fused = (has_bn and training and fused_finalize()
                 and (dp is None or world == 1 or getattr(dp, "_peer", None) is not None))
        peer = _peer_args(dp)

        def bn_finalize(i, count, C_, affine=True, centered=False, nparts=0):
            """sums[i] -> scale/shift/mean/rstd of BN layer i (batch or running statistics).
            affine=False gives the plain normalisation (scale = rstd, shift = -mean*rstd); centered=True means the
            statistics were taken on x - running_mean (the single-pass input kernel).  nparts > 0: the producer left
            nparts partial rows at the start of `scratch` (fused finaliser)."""
            bnm = getattr(mod, "bn%d" % i)
            gam = P["bn%d.weight" % i].data_ptr() if affine else None
            bet = P["bn%d.bias" % i].data_ptr() if affine else None
            ctr = bnm.running_mean.data_ptr() if centered else None
            if nparts:
                L.call("dcue_bn_stats_finalize", scratch, int(nparts), float(count * world), C_, gam, bet,
                       bnm.running_mean.data_ptr(), bnm.running_var.data_ptr(), bnm.num_batches_tracked.data_ptr(), BN_MOMENTUM,
                       BN_EPS, ctr, *peer, ws.ticket.data_ptr(), ws.sums[i].data_ptr(), ws.bnp[i, 0].data_ptr(), ws.bnp[i, 1].data_ptr(),
                       ws.bnp[i, 2].data_ptr(), ws.bnp[i, 3].data_ptr(), st)
                return
            if training and dp is not None:
                dp.all_reduce_sum(ws.sums[i])
            L.call("dcue_bn_finalize", ws.sums[i].data_ptr(), float(count * world), C_, gam, bet, bnm.running_mean.data_ptr(),
                   bnm.running_var.data_ptr(), bnm.num_batches_tracked.data_ptr(), BN_MOMENTUM, BN_EPS, int(training), ctr,
                   ws.bnp[i, 0].data_ptr(), ws.bnp[i, 1].data_ptr(), ws.bnp[i, 2].data_ptr(), ws.bnp[i, 3].data_ptr(), st)

        pos_p, neg_p = pos.data_ptr(), (None if neg is None else neg.data_ptr())
        # ---- input -> layer1 operand panel.  BatchNorm towers: ONE pass over the fp32 input writes
        # u = x - bn0.running_mean as the fp16 panel and accumulates the batch statistics of u; the exact
        # normalisation x-hat = rstd*u + shift, bn0's gamma and beta are all folded into layer1 (packed weights
        # W*gamma*rstd, border-aware bias), which also lets the backward skip layer1's data gradient.
        g0 = geo[0]
        off_p = None if (src is None or off is None) else off.data_ptr()
        if has_bn:
            rm0 = mod.bn0.running_mean.data_ptr()
            sums0 = None if fused else ws.sums[0].data_ptr()      # None: partial rows stay in scratch for the fused finaliser
            if src is not None:
                L.call("dcue_ncl_center_pack_stats_indexed", pos_p, n_songs, T, idx.data_ptr(), off_p, S, C, frames, err.data_ptr(),
                       rm0, ws.X[0].base, ws.X[0].panel_rows, g0["Lp"], g0["pad"], fmt, sums0, scratch, nscr, st)
            else:
                L.call("dcue_ncl_center_pack_stats", pos_p, S_pos, neg_p, S_neg, C, frames, rm0, ws.X[0].base, ws.X[0].panel_rows,
                       g0["Lp"], g0["pad"], fmt, sums0, scratch, nscr, st)
            # bnp[0] = (rstd, shift, mean_u, rstd)
            bn_finalize(0, S * frames, C, affine=False, centered=True,
                        nparts=L.query("dcue_ncl_center_pack_stats_nparts", S, C) if fused else 0)
            L.call("dcue_conv_tap_bias", P["layer1.weight"].data_ptr(), H, 128, g0["k"], P["bn0.bias"].data_ptr(),
                   P["bn0.weight"].data_ptr(), ws.bnp[0, 1].data_ptr(), ws.tapb.data_ptr(), st)
        elif src is not None:
            L.call("dcue_ncl_pack_indexed", pos_p, n_songs, T, idx.data_ptr(), off_p, S, C, frames, err.data_ptr(), None, None,
                   ws.X[0].base, ws.X[0].panel_rows, g0["Lp"], g0["pad"], fmt, st)
        else:
            L.call("dcue_ncl_pack", pos_p, S_pos, neg_p, S_neg, C, frames, None, None, ws.X[0].base, ws.X[0].panel_rows,
                   g0["Lp"], g0["pad"], fmt, st)
        # ---- layer1..4: conv + pool + relu (+ BN statistics) -> affine -> next operand panel
        for i, g in enumerate(geo, start=1):
            Wt, bt = P["layer%d.weight" % i], P["layer%d.bias" % i]
            fold = has_bn and i == 1
            L.call("dcue_pack_conv_weight", Wt.data_ptr(), H, 128, g["k"], 0, fmt,
                   P["bn0.weight"].data_ptr() if fold else None, ws.bnp[0, 0].data_ptr() if fold else None,
                   ws.wp[i - 1].data_ptr(), st)
            want_stats = has_bn and training
            if fused:
                L.call("dcue_conv_pool_fwd_parts", impl, ws.X[i - 1].base, ws.X[i - 1].panel_rows, fmt, ws.wp[i - 1].data_ptr(),
                       bt.data_ptr(), ws.tapb.data_ptr() if fold else None, S, g["Lp"], g["Lin"], g["pad"], g["P"], g["pool"],
                       g["k"], 128, H, ws.z[i - 1].data_ptr(), ws.code[i - 1].data_ptr(), scratch, nscr, st)
            else:
                L.call("dcue_conv_pool_fwd", impl, ws.X[i - 1].base, ws.X[i - 1].panel_rows, fmt, ws.wp[i - 1].data_ptr(),
                       bt.data_ptr(), ws.tapb.data_ptr() if fold else None, S, g["Lp"], g["Lin"], g["pad"], g["P"], g["pool"],
                       g["k"], 128, H, ws.z[i - 1].data_ptr(),
                       ws.code[i - 1].data_ptr(), ws.sums[i].data_ptr() if want_stats else None, scratch, nscr, st)
            if has_bn:
                bn_finalize(i, S * g["P"], H, nparts=L.query("dcue_conv_pool_fwd_nparts", impl, S, g["Lp"]) if fused else 0)
                sc, sh = ws.bnp[i, 0].data_ptr(), ws.bnp[i, 1].data_ptr()
            else:
                sc = sh = None
            tp = ws.fc_in[:, (i - 1) * H:].data_ptr() if res else None
            ldtp = ws.fc_in.shape[1]
            if i < 4:
                nx = geo[i]
                L.call("dcue_affine_pack", ws.z[i - 1].data_ptr(), S, g["P"], H, sc, sh, ws.X[i].base, ws.X[i].panel_rows,
                       nx["Lp"], nx["pad"], fmt, None, tp, ldtp, st)
            else:
                L.call("dcue_affine_pack", ws.z[3].data_ptr(), S, 1, H, sc, sh, None, 0, 1, 0, fmt, ws.y4.data_ptr(), tp,
                       ldtp, st)
        # ---- layer5 (k=1 conv == linear) + relu (+bn5), fc
        L.call(tower_linear("dcue_linear_fwd"), ws.y4.data_ptr(), H, P["layer5.weight"].data_ptr(), P["layer5.bias"].data_ptr(), S, H, F,
               1, ws.z5.data_ptr(), F, st)
        y5 = ws.fc_in[:, 4 * H:] if res else ws.fc_in
        if has_bn:
            if training:  # sum z, sum z^2 via the backward-reduce kernel with mean=0, rstd=1
                L.call("dcue_bn_bwd_reduce", ws.z5.data_ptr(), F, None, 0, ws.z5.data_ptr(), ws.zero128.data_ptr(),
                       ws.one128.data_ptr(), S, 1, F, None if fused else ws.sums[5].data_ptr(), None, None, None, scratch, nscr, st)
            bn_finalize(5, S, F, nparts=L.query("dcue_bn_bwd_reduce_nparts", S, 1) if fused else 0)
            sc, sh = ws.bnp[5, 0].data_ptr(), ws.bnp[5, 1].data_ptr()
        else:
            sc = sh = None
        # strided affine copy of z5 into its slot of the fc input (time mean over P=1 rows)
        L.call("dcue_affine_pack", ws.z5.data_ptr(), S, 1, F, sc, sh, None, 0, 1, 0, fmt, None, y5.data_ptr(),
               ws.fc_in.shape[1], st)
        out = torch.empty(S, F, dtype=torch.float32, device=dev)
        Kfc = ws.fc_in.shape[1]
        L.call(tower_linear("dcue_linear_fwd"), ws.fc_in.data_ptr(), Kfc, P["fc.weight"].data_ptr(), P["fc.bias"].data_ptr(), S, Kfc, F,
               0, out.data_ptr(), F, st)

        needs_bwd = any(ctx.needs_input_grad)
        if needs_bwd:
            ctx.ws, ctx.mod, ctx.training = ws, mod, training
            ctx.pos, ctx.neg, ctx.dims = pos, neg, (S_pos, S_neg, C, frames)
            ctx.impl, ctx.fmt, ctx.world = impl, fmt, world
            ctx.save_for_backward(*params)
        else:
            _release(ws)
        return out

    @staticmethod
    def backward(ctx, gout):
        ws, mod, training = ctx.ws, ctx.mod, ctx.training
        if ws is None:
            raise RuntimeError("SongTowerFn backward called twice (workspace already released)")
        has_bn, res = mod._has_bn, mod._res
        H, F = mod.hidden_size, mod.output_size
        S_pos, S_neg, C, frames = ctx.dims
        S = S_pos + S_neg
        geo, impl, fmt, world = ws.geo, ctx.impl, ctx.fmt, ctx.world
        dp = mod._dp
        params = ctx.saved_tensors
        P = dict(zip(mod._param_names, params))
        dev = gout.device
        st = L.stream()
        scratch, nscr = ws.scratch.data_ptr(), ws.scratch.numel()
        b = ws.bwd()
        f32 = dict(dtype=torch.float32, device=dev)
        gout = gout.contiguous()
        grads = {}
        Kfc = ws.fc_in.shape[1]
        # tcgen05 kind::f16 wants both MMA operands in ONE 16-bit format, so the gradient operand dY
        # uses the forward operand format; a per-layer power-of-two scale keeps fp16 in range.
        gfmt = fmt
        bn_train = has_bn and training

        fused_fin = fused_finalize() and (dp is None or world == 1 or getattr(dp, "_peer", None) is not None)

        def bn_sums(i, dy_ptr, lddy, dtp_ptr, lddtp, z, P_, C_, want_scale=True, pre_nparts=0):
            """Per-channel reductions over the gradient entering layer i's BN (+ DP all-reduce):
            dgamma/dbeta when BN uses batch statistics, and max|dy| for the 16-bit gradient scale."""
            mean_p = ws.bnp[i, 2].data_ptr() if bn_train else ws.zero128.data_ptr()
            rstd_p = ws.bnp[i, 3].data_ptr() if bn_train else ws.one128.data_ptr()
            gw = gb = None
            if bn_train:
                gw, gb = torch.empty(C_, **f32), torch.empty(C_, **f32)
            if fused_fin:
                # one sweep leaves per-block partials in scratch, ONE kernel reduces them, all-reduces the sums over NVLink
                # (data parallel), emits dgamma / dbeta in fp32 and the power-of-two scale of the 16-bit gradient operand
                if dy_ptr is not None:      # None: the producing dgrad already left the partial rows in scratch
                    L.call("dcue_bn_bwd_reduce", dy_ptr, lddy, dtp_ptr, lddtp, z.data_ptr(), mean_p, rstd_p, S, P_, C_, None, None,
                           None, None, scratch, nscr, st)
                    nparts = L.query("dcue_bn_bwd_reduce_nparts", S, P_)
                else:
                    nparts = pre_nparts
                peer = _peer_args(dp) if bn_train else (None, None, None, 0, 1)
                L.call("dcue_bn_bwd_finalize", scratch, nparts, C_,
                       ws.bnp[i, 0].data_ptr() if has_bn else None, float(S * P_ * world) if bn_train else 0.0, *peer,
                       ws.ticket.data_ptr(), b["dsums"][i].data_ptr(), L.ptr(gb), L.ptr(gw),
                       b["amax"][i:].data_ptr() if want_scale else None, b["gscale"][i].data_ptr() if want_scale else None, st)
                if bn_train:
                    grads["bn%d.weight" % i], grads["bn%d.bias" % i] = gw, gb
                elif has_bn:
                    grads["bn%d.weight" % i] = grads["bn%d.bias" % i] = None  # eval-mode backward: constants
                return
            direct = bn_train and dp is None  # single GPU: the reducer emits dgamma / dbeta in fp32 directly
            L.call("dcue_bn_bwd_reduce", dy_ptr, lddy, dtp_ptr, lddtp, z.data_ptr(), mean_p, rstd_p, S, P_, C_,
                   b["dsums"][i].data_ptr(), b["amax"][i:].data_ptr() if want_scale else None,
                   gb.data_ptr() if direct else None, gw.data_ptr() if direct else None, scratch, nscr, st)
            if bn_train:
                if dp is not None:
                    dp.all_reduce_sum(b["dsums"][i])
                    L.call("dcue_cvt_f64_f32", b["dsums"][i][C_:].data_ptr(), C_, 1.0, gw.data_ptr(), st)
                    L.call("dcue_cvt_f64_f32", b["dsums"][i].data_ptr(), C_, 1.0, gb.data_ptr(), st)
                grads["bn%d.weight" % i], grads["bn%d.bias" % i] = gw, gb
            elif has_bn:
                grads["bn%d.weight" % i] = grads["bn%d.bias" % i] = None  # eval-mode backward: constants
            if want_scale:
                L.call("dcue_grad_scale", b["amax"][i:].data_ptr(), ws.bnp[i, 0].data_ptr() if has_bn else None, C_,
                       float(S * P_ * world) if bn_train else 0.0, b["gscale"][i].data_ptr(), st)

        # ---- fc
        gW, gb_ = torch.empty(F, Kfc, **f32), torch.empty(F, **f32)
        L.call(tower_linear("dcue_linear_wgrad"), gout.data_ptr(), F, ws.fc_in.data_ptr(), Kfc, S, Kfc, F, gW.data_ptr(), gb_.data_ptr(),
               scratch, nscr, st)
        grads["fc.weight"], grads["fc.bias"] = gW, gb_
        dfc = torch.empty(S, Kfc, **f32)
        L.call(tower_linear("dcue_linear_dgrad"), gout.data_ptr(), F, P["fc.weight"].data_ptr(), S, Kfc, F, None, 0, dfc.data_ptr(), Kfc, st)
        dy5 = dfc[:, 4 * H:] if res else dfc
        # ---- bn5 + relu5 -> dz5 ; layer5
        dz5 = torch.empty(S, F, **f32)
        if has_bn:
            bn_sums(5, dy5.data_ptr(), Kfc, None, 0, ws.z5, 1, F, want_scale=False)
        L.call("dcue_bn_relu_unpool_bwd", dy5.data_ptr(), Kfc, None, 0, ws.z5.data_ptr(), None,
               ws.bnp[5, 0].data_ptr() if has_bn else None, ws.bnp[5, 2].data_ptr() if has_bn else None,
               ws.bnp[5, 3].data_ptr() if has_bn else None, b["dsums"][5].data_ptr() if bn_train else None,
               float(S * world), S, 1, F, 1, 1, None, 0, gfmt, None, dz5.data_ptr(), None, None, scratch, nscr, st)
        gW5, gb5 = torch.empty(F, H, 1, **f32), torch.empty(F, **f32)
        L.call(tower_linear("dcue_linear_wgrad"), dz5.data_ptr(), F, ws.y4.data_ptr(), H, S, H, F, gW5.data_ptr(), gb5.data_ptr(), scratch,
               nscr, st)
        grads["layer5.weight"], grads["layer5.bias"] = gW5, gb5
        dy = torch.empty(S, H, **f32)  # gradient w.r.t. the stage-4 BN output
        L.call(tower_linear("dcue_linear_dgrad"), dz5.data_ptr(), F, P["layer5.weight"].data_ptr(), S, H, F, None, 0, dy.data_ptr(), H, st)
        # ---- stages 4..1
        stats_in_scratch = 0
        for i in range(4, 0, -1):
            g = geo[i - 1]
            dtp = dfc[:, (i - 1) * H:].data_ptr() if res else None
            if stats_in_scratch:     # the dgrad of the stage above took the reductions in its epilogue
                bn_sums(i, None, H, dtp, Kfc, ws.z[i - 1], g["P"], H, pre_nparts=stats_in_scratch)
                stats_in_scratch = 0
            else:
                bn_sums(i, dy.data_ptr(), H, dtp, Kfc, ws.z[i - 1], g["P"], H)
            gsc = b["gscale"][i].data_ptr()
            gb_i = torch.empty(H, **f32)
            gW_i = torch.empty(H, 128, g["k"], **f32)
            bn_args = (ws.bnp[i, 0].data_ptr() if has_bn else None, ws.bnp[i, 2].data_ptr() if has_bn else None,
                       ws.bnp[i, 3].data_ptr() if has_bn else None, b["dsums"][i].data_ptr() if bn_train else None,
                       float(S * g["P"] * world))
            # layer 1 needs no data gradient: its dY operand is built inside the weight-gradient kernel
            fused = i == 1 and impl == L.IMPL_TC and g["pool"] == 4 and g["k"] == 4 and fused_wgrad()
            dYp = None if fused else ws.dy_panel(i - 1)
            if fused:
                L.call("dcue_conv_wgrad_unpool", dy.data_ptr(), H, dtp, Kfc, ws.z[i - 1].data_ptr(), ws.code[i - 1].data_ptr(),
                       *bn_args, S, g["P"], g["pool"], g["Lp"], ws.X[i - 1].base, ws.X[i - 1].panel_rows, fmt, g["k"], 128, H,
                       gsc, gW_i.data_ptr(), b["bsum"].data_ptr(), gb_i.data_ptr(), scratch, nscr, st)
            else:
                L.call("dcue_bn_relu_unpool_bwd", dy.data_ptr(), H, dtp, Kfc, ws.z[i - 1].data_ptr(), ws.code[i - 1].data_ptr(),
                       *bn_args, S, g["P"], H, g["pool"], g["Lp"], dYp.base, dYp.panel_rows, gfmt, gsc, None,
                       b["bsum"].data_ptr(), gb_i.data_ptr(), scratch, nscr, st)
                L.call("dcue_conv_wgrad", impl, dYp.base, dYp.panel_rows, gfmt, ws.X[i - 1].base, ws.X[i - 1].panel_rows, fmt,
                       S * g["Lp"], g["k"], 128, H, gsc, gW_i.data_ptr(), scratch, nscr, st)
            if i == 1 and has_bn:
                # gW_i is G = sum dY * xhat.  bn0 was folded into this conv: its gradients and the true
                # weight gradient follow from G and a few border row sums -- no data gradient needed.
                k_, pad_, lin_ = g["k"], g["pad"], g["Lin"]
                brow = list(range(pad_)) + list(range(lin_ + pad_ - k_ + 1, lin_ + 2 * pad_ - k_ + 1))
                brow = (brow + [-1, -1, -1, -1])[:4]
                if fused:
                    L.call("dcue_border_row_sums", dy.data_ptr(), H, dtp, Kfc, ws.z[0].data_ptr(), ws.code[0].data_ptr(), *bn_args,
                           S, g["P"], H, g["pool"], brow[0], brow[1], brow[2], brow[3], b["E"].data_ptr(), st)
                else:
                    L.call("dcue_panel_row_sums", dYp.base, dYp.panel_rows, gfmt, S, g["Lp"], brow[0], brow[1], brow[2], brow[3],
                           gsc, b["E"].data_ptr(), st)
                dW1, dg0, db0 = torch.empty(H, 128, k_, **f32), torch.empty(128, **f32), torch.empty(128, **f32)
                L.call("dcue_bn_fold_grads", gW_i.data_ptr(), P["layer1.weight"].data_ptr(), P["bn0.weight"].data_ptr(),
                       P["bn0.bias"].data_ptr(), ws.bnp[0, 0].data_ptr(), ws.bnp[0, 1].data_ptr(), gb_i.data_ptr(),
                       b["E"].data_ptr(), brow[0], brow[1], brow[2], brow[3], H, 128,
                       k_, pad_, lin_, dW1.data_ptr(), dg0.data_ptr(), db0.data_ptr(), st)
                gW_i = dW1
                grads["bn0.weight"], grads["bn0.bias"] = dg0, db0
            grads["layer%d.weight" % i], grads["layer%d.bias" % i] = gW_i, gb_i
            if i > 1:
                L.call("dcue_pack_conv_weight", P["layer%d.weight" % i].data_ptr(), H, 128, g["k"], 1, fmt, None, None,
                       b["wpd"][i - 1].data_ptr(), st)
                dx = b["dx"][i - 1]
                if fused_fin and impl == L.IMPL_TC and dgrad_stats():
                    # dx is the gradient entering stage i-1's BatchNorm: its backward reductions are taken in this epilogue
                    j = i - 1
                    mean_j = ws.bnp[j, 2].data_ptr() if bn_train else ws.zero128.data_ptr()
                    rstd_j = ws.bnp[j, 3].data_ptr() if bn_train else ws.one128.data_ptr()
                    dtp_j = dfc[:, (j - 1) * H:].data_ptr() if res else None
                    L.call("dcue_conv_dgrad_stats", dYp.base, dYp.panel_rows, gfmt, b["wpd"][i - 1].data_ptr(), fmt, S, g["Lp"],
                           g["Lin"], g["pad"], g["k"], 128, H, gsc, dx.data_ptr(), ws.z[j - 1].data_ptr(), mean_j, rstd_j,
                           dtp_j, Kfc, scratch, nscr, st)
                    stats_in_scratch = L.query("dcue_conv_pool_fwd_nparts", impl, S, g["Lp"])
                else:
                    L.call("dcue_conv_dgrad", impl, dYp.base, dYp.panel_rows, gfmt, b["wpd"][i - 1].data_ptr(), fmt, S, g["Lp"],
                           g["Lin"], g["pad"], g["k"], 128, H, gsc, dx.data_ptr(), scratch, nscr, st)
                dy = dx
        ctx.ws = None
        _release(ws)
        join_backward_side()        # the user tower's backward (side stream) ends inside this backward pass
        return (None, None, None, None, None) + tuple(grads.get(n) for n in mod._param_names)


class UserTowerFn(torch.autograd.Function):
    """u_f = linear2(relu(linear1(relu(table[idx]))))  (userembedding.py:40-44)."""

    @staticmethod
    def forward(ctx, idx, table, w1, b1, w2, b2, err, dp=None, overlap_backward=False):
        ctx.dp = dp
        ctx.overlap = bool(overlap_backward)
        ctx.biases = (b1, b2)
        join_backward_side()
        if not table.is_cuda:
            raise RuntimeError("the DCUE B200 path has no CPU fallback: move the model to a CUDA device")
        if idx.dtype != torch.int64:
            raise TypeError("user indices must be int64")
        shape = idx.shape
        idx = idx.to(table.device).contiguous().view(-1)
        B, (U, E), F = idx.numel(), table.shape, w2.shape[0]
        dev, st = table.device, L.stream()
        f32 = dict(dtype=torch.float32, device=dev)
        h0, h1, out = torch.empty(B, E, **f32), torch.empty(B, E, **f32), torch.empty(B, F, **f32)
        L.call("dcue_gather_relu_fwd", table.data_ptr(), idx.data_ptr(), B, U, E, h0.data_ptr(), None, err.data_ptr(), st)
        L.call("dcue_linear_fwd", h0.data_ptr(), E, w1.data_ptr(), b1.data_ptr(), B, E, E, 1, h1.data_ptr(), E, st)
        L.call("dcue_linear_fwd", h1.data_ptr(), E, w2.data_ptr(), b2.data_ptr(), B, E, F, 0, out.data_ptr(), F, st)
        # out-of-range indices set `err` and poison their rows with NaN; the flag is read lazily
        # (UserEmbeddings.raise_if_index_error) so that the forward never synchronises with the host
        ctx.save_for_backward(idx, table, w1, w2, h0, h1)
        return out.view(*shape, F)

    @staticmethod
    def backward(ctx, gout):
        idx, table, w1, w2, h0, h1 = ctx.saved_tensors
        fork = (ctx.overlap and table.is_cuda and os.environ.get("DCUE_USER_BWD_STREAM", "1") != "0"
                and all(t.grad is None for t in (table, w1, w2) + tuple(ctx.biases)))
        if not fork:
            return UserTowerFn._backward(ctx, gout)
        cur = torch.cuda.current_stream()
        side = _bwd_side_stream(table.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            out = UserTowerFn._backward(ctx, gout)
        for t in (gout, idx, h0, h1):
            t.record_stream(side)             # allocated on the main stream, read on the side stream
        for t in out:
            if t is not None:
                t.record_stream(cur)          # produced on the side stream, consumed on the main stream after the join
        _BWD_SIDE["pending"] = (side, (gout, idx, h0, h1))    # inputs only: a second reference to a gradient makes autograd clone it
        return out

    @staticmethod
    def _backward(ctx, gout):
        idx, table, w1, w2, h0, h1 = ctx.saved_tensors
        B, (U, E), F = idx.numel(), table.shape, w2.shape[0]
        dev, st = table.device, L.stream()
        f32 = dict(dtype=torch.float32, device=dev)
        gout = gout.contiguous().view(B, F)
        nscr = max(L.query("dcue_linear_wgrad_ws_bytes", B, E, F), L.query("dcue_linear_wgrad_ws_bytes", B, E, E))
        scratch = torch.empty(nscr, dtype=torch.uint8, device=dev)
        gw2, gb2 = torch.empty(F, E, **f32), torch.empty(F, **f32)
        L.call("dcue_linear_wgrad", gout.data_ptr(), F, h1.data_ptr(), E, B, E, F, gw2.data_ptr(), gb2.data_ptr(),
               scratch.data_ptr(), nscr, st)
        dh1 = torch.empty(B, E, **f32)
        L.call("dcue_linear_dgrad", gout.data_ptr(), F, w2.data_ptr(), B, E, F, h1.data_ptr(), E, dh1.data_ptr(), E, st)
        gw1, gb1 = torch.empty(E, E, **f32), torch.empty(E, **f32)
        L.call("dcue_linear_wgrad", dh1.data_ptr(), E, h0.data_ptr(), E, B, E, E, gw1.data_ptr(), gb1.data_ptr(),
               scratch.data_ptr(), nscr, st)
        gtable = None
        dp = ctx.dp
        if ctx.needs_input_grad[1] and dp is not None and dp.world_size > 1:
            # data parallel: exchange the B ReLU-masked gradient rows (+ their indices) instead of all-reducing the
            # dense [U,E] gradient (1.2 MB per rank instead of 24 MB at cfg3); every rank then segment-sums the same
            # world*B rows in the same order, so the dense gradient is already the global sum and identical everywhere
            xch = dp.table_exchange(B, E, dev) if hasattr(dp, "table_exchange") else None
            if xch is not None:
                # NVLink peer memory: the masked gradient rows are written straight into this rank's exchange slot, one
                # single-CTA kernel publishes the indices + barriers, and every rank sums all world*B rows in (rank, position)
                # order reading them from the peers' slots -- no NCCL all-gather, no staging copies.
                # The whole exchange (100 us on 8 GPUs, on the critical path when it ran at the end of backward) goes to a
                # SIDE STREAM: this backward runs before the song tower's (DCUENet evaluates the user tower last), so the
                # exchange overlaps ~1.2 ms of tower kernels; DataParallelDCUE.reduce_gradients() joins the stream.  Only when
                # autograd will ASSIGN the table gradient (table.grad is None: GraphedTrainStep, zero_grad(set_to_none=True)):
                # an accumulating `grad += gtable` would run on the main stream without waiting for the side stream.
                side = dp.exchange_stream(table) if hasattr(dp, "exchange_stream") else None
                cur = torch.cuda.current_stream()
                if side is not None:
                    side.wait_stream(cur)
                with torch.cuda.stream(side if side is not None else cur):
                    xch.barrier(0)          # every peer has finished reading the previous step's rows
                    L.call("dcue_linear_dgrad", dh1.data_ptr(), E, w1.data_ptr(), B, E, E, h0.data_ptr(), E, xch.rows.data_ptr(), E,
                           L.stream())
                    gtable = xch.scatter_add(xch.exchange_indices(idx), B, 0, U)
                if side is not None:
                    for t in (dh1, h0, idx, w1):
                        t.record_stream(side)
                    gtable.record_stream(cur)
                    # NOT gtable itself: an extra reference makes AccumulateGrad clone the gradient (on the main stream,
                    # before the side stream has written it) instead of adopting the tensor
                    dp.exchange_pending(side, (dh1, h0, idx, w1))
            else:
                drows = torch.empty(B, E, **f32)
                L.call("dcue_linear_dgrad", dh1.data_ptr(), E, w1.data_ptr(), B, E, E, h0.data_ptr(), E, drows.data_ptr(), E, st)
                gtable = scatter_rows(dp.all_gather_rows(idx), dp.all_gather_rows(drows), U)
        elif ctx.needs_input_grad[1]:
            dh0 = torch.empty(B, E, **f32)  # ReLU mask of the gather is applied in the scatter kernel
            L.call("dcue_linear_dgrad", dh1.data_ptr(), E, w1.data_ptr(), B, E, E, None, 0, dh0.data_ptr(), E, st)
            gtable = scatter_rows(idx, dh0, U, mask=h0)     # dense gradient, like nn.Embedding(sparse=False)
        return None, gtable, gw1, gb1, gw2, gb2, None, None, None


class UserMLPFn(torch.autograd.Function):
    """u_f = linear2(relu(linear1(relu(rows)))) on pre-gathered raw table rows [B,E] (row-sharded table path:
    the gather happened on the owning ranks).  Gradient w.r.t. rows is already ReLU-masked."""

    @staticmethod
    def forward(ctx, rows, w1, b1, w2, b2):
        rows = rows.contiguous()
        B, E = rows.shape
        F = w2.shape[0]
        dev, st = rows.device, L.stream()
        f32 = dict(dtype=torch.float32, device=dev)
        h0, h1, out = torch.empty(B, E, **f32), torch.empty(B, E, **f32), torch.empty(B, F, **f32)
        iota = torch.arange(B, dtype=torch.int64, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        L.call("dcue_gather_relu_fwd", rows.data_ptr(), iota.data_ptr(), B, B, E, h0.data_ptr(), None, err.data_ptr(), st)
        L.call("dcue_linear_fwd", h0.data_ptr(), E, w1.data_ptr(), b1.data_ptr(), B, E, E, 1, h1.data_ptr(), E, st)
        L.call("dcue_linear_fwd", h1.data_ptr(), E, w2.data_ptr(), b2.data_ptr(), B, E, F, 0, out.data_ptr(), F, st)
        ctx.save_for_backward(w1, w2, h0, h1)
        return out

    @staticmethod
    def backward(ctx, gout):
        w1, w2, h0, h1 = ctx.saved_tensors
        B, E = h0.shape
        F = w2.shape[0]
        dev, st = h0.device, L.stream()
        f32 = dict(dtype=torch.float32, device=dev)
        gout = gout.contiguous()
        nscr = max(L.query("dcue_linear_wgrad_ws_bytes", B, E, F), L.query("dcue_linear_wgrad_ws_bytes", B, E, E))
        scratch = torch.empty(nscr, dtype=torch.uint8, device=dev)
        gw2, gb2 = torch.empty(F, E, **f32), torch.empty(F, **f32)
        L.call("dcue_linear_wgrad", gout.data_ptr(), F, h1.data_ptr(), E, B, E, F, gw2.data_ptr(), gb2.data_ptr(),
               scratch.data_ptr(), nscr, st)
        dh1 = torch.empty(B, E, **f32)
        L.call("dcue_linear_dgrad", gout.data_ptr(), F, w2.data_ptr(), B, E, F, h1.data_ptr(), E, dh1.data_ptr(), E, st)
        gw1, gb1 = torch.empty(E, E, **f32), torch.empty(E, **f32)
        L.call("dcue_linear_wgrad", dh1.data_ptr(), E, h0.data_ptr(), E, B, E, E, gw1.data_ptr(), gb1.data_ptr(),
               scratch.data_ptr(), nscr, st)
        drows = torch.empty(B, E, **f32)  # masked by the gather's ReLU (h0 > 0)
        L.call("dcue_linear_dgrad", dh1.data_ptr(), E, w1.data_ptr(), B, E, E, h0.data_ptr(), E, drows.data_ptr(), E, st)
        return drows, gw1, gb1, gw2, gb2


def gather_rows(table, idx):
    """raw table rows [B,E] with the gather kernel (idx int64 within [0, table.shape[0]))."""
    B, (U, E) = idx.numel(), table.shape
    dev = table.device
    relu_out = torch.empty(B, E, dtype=torch.float32, device=dev)
    raw = torch.empty(B, E, dtype=torch.float32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call("dcue_gather_relu_fwd", table.data_ptr(), idx.contiguous().data_ptr(), B, U, E, relu_out.data_ptr(), raw.data_ptr(),
           err.data_ptr(), L.stream())
    return raw


def scatter_rows(idx, grad_rows, n_rows, mask=None):
    """Dense [n_rows,E] sum of grad_rows by idx, deterministic (duplicates are added in position order); entries whose
    index is outside [0, n_rows) -- the 'not mine' sentinel of the sharded table, or a bad user index already flagged by
    the forward gather -- are dropped.  mask (optional, [B,E]): ReLU output of the gather, gradient passes where > 0.
    Step-sized batches take the sort-free single-launch kernel; larger ones sort by row first."""
    B, E = grad_rows.shape
    dev = grad_rows.device
    out = torch.zeros(n_rows, E, dtype=torch.float32, device=dev)
    if B == 0:
        return out
    st = L.stream()
    idx, grad_rows = idx.contiguous(), grad_rows.contiguous()
    mask = None if mask is None else mask.contiguous()
    if B <= L.lib().dcue_scatter_direct_max():
        L.call("dcue_scatter_add_rows", grad_rows.data_ptr(), L.ptr(mask), idx.data_ptr(), B, n_rows, E, out.data_ptr(), st)
        return out
    sidx = torch.empty(B, dtype=torch.int64, device=dev)
    spos = torch.empty(B, dtype=torch.int32, device=dev)
    nscr = L.query("dcue_sort_ws_bytes", B)
    scratch = torch.empty(nscr, dtype=torch.uint8, device=dev)
    # out-of-range keys: the sort orders by the low bits only, which is fine -- equal keys stay adjacent and the scatter
    # kernel skips every row outside [0, n_rows)
    L.call("dcue_sort_indices", idx.data_ptr(), B, n_rows + 1, sidx.data_ptr(), spos.data_ptr(), scratch.data_ptr(), nscr, st)
    if mask is None:
        mask = torch.ones_like(grad_rows)
    L.call("dcue_scatter_add_bwd", grad_rows.data_ptr(), mask.data_ptr(), sidx.data_ptr(), spos.data_ptr(), B, n_rows, E,
           out.data_ptr(), st)
    return out


class ScoreFn(torch.autograd.Function):
    """scores[b,n] = cos(u_b, pos_b) - cos(u_b, neg_bn)  (dcue.py:93-106); feats = [pos; neg]."""

    @staticmethod
    def forward(ctx, u_f, feats, B, N):
        u_f, feats = u_f.contiguous(), feats.contiguous()
        F = u_f.shape[1]
        scores = torch.empty(B, N, dtype=torch.float32, device=u_f.device)
        if B * N > 0:
            L.call("dcue_score_fwd",     u_f.data_ptr(), feats.data_ptr(), B, N, F, COS_EPS, scores.data_ptr(), L.stream())
        ctx.save_for_backward(u_f, feats)
        ctx.dims = (B, N, F)
        return scores

    @staticmethod
    def backward(ctx, gs):
        u_f, feats = ctx.saved_tensors
        B, N, F = ctx.dims
        du, df = torch.empty_like(u_f), torch.empty_like(feats)
        L.call("dcue_score_bwd", u_f.data_ptr(), feats.data_ptr(), gs.contiguous().data_ptr(), B, N, F, COS_EPS,
               du.data_ptr(), df.data_ptr(), L.stream())
        return du, df, None, None


class HingeScoreFn(torch.autograd.Function):
    """Fused scores + max-margin hinge loss + analytic backward in ONE kernel launch.
    Returns (loss_rows[B], scores[B,N]); mean over the GLOBAL batch is folded into the gradient."""

    @staticmethod
    def forward(ctx, u_f, feats, B, N, margin, batch_total):
        u_f, feats = u_f.contiguous(), feats.contiguous()
        F = u_f.shape[1]
        dev = u_f.device
        scores = torch.empty(B, N, dtype=torch.float32, device=dev)
        loss_rows = torch.empty(B, dtype=torch.float32, device=dev)
        du, df = torch.empty_like(u_f), torch.empty_like(feats)
        L.call("dcue_score_hinge_fwdbwd", u_f.data_ptr(), feats.data_ptr(), B, N, F, COS_EPS, float(margin), int(batch_total),
               scores.data_ptr(), loss_rows.data_ptr(), du.data_ptr(), df.data_ptr(), L.stream())
        ctx.save_for_backward(du, df)
        ctx.batch_total = batch_total
        ctx.mark_non_differentiable(scores)
        return loss_rows, scores

    @staticmethod
    def backward(ctx, g_rows, _g_scores):
        # du/df hold d(sum_b loss_rows / batch_total); the caller's loss is loss_rows.sum()/batch_total,
        # i.e. g_rows == 1/batch_total for every row -> rescale by g_rows * batch_total (a scalar).
        du, df = ctx.saved_tensors
        scale = g_rows.reshape(-1)[:1] * float(ctx.batch_total)
        return du * scale, df * scale, None, None, None, None


class HingeLossFn(torch.autograd.Function):
    """HingeScoreFn + the mean over the (global) batch as ONE differentiable op: forward = the fused score/hinge kernel and a
    one-block sum (scalar loss on the device), backward = one kernel scaling the stored gradients by the incoming gradient.
    Replaces `loss_rows.sum() / total` and its autograd chain (8 ATen launches per step in the round-2 timeline).
    Returns (loss [], scores [B, N])."""

    @staticmethod
    def forward(ctx, u_f, feats, B, N, margin, batch_total):
        u_f, feats = u_f.contiguous(), feats.contiguous()
        F = u_f.shape[1]
        dev = u_f.device
        scores = torch.empty(B, N, dtype=torch.float32, device=dev)
        loss_rows = torch.empty(B, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        du, df = torch.empty_like(u_f), torch.empty_like(feats)
        L.call("dcue_score_hinge_fwdbwd", u_f.data_ptr(), feats.data_ptr(), B, N, F, COS_EPS, float(margin), int(batch_total),
               scores.data_ptr(), loss_rows.data_ptr(), du.data_ptr(), df.data_ptr(), L.stream())
        L.call("dcue_loss_mean", loss_rows.data_ptr(), B, int(batch_total), loss.data_ptr(), L.stream())
        ctx.save_for_backward(du, df)
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, g, _g_scores):
        du, df = ctx.saved_tensors          # d(loss)/d(u_f), d(loss)/d(feats) for an incoming gradient of 1
        g = g.to(torch.float32).contiguous()
        gu, gf = torch.empty_like(du), torch.empty_like(df)
        L.call("dcue_scale_pair", du.data_ptr(), du.numel(), df.data_ptr(), df.numel(), g.data_ptr(), gu.data_ptr(), gf.data_ptr(),
               L.stream())
        return gu, gf, None, None, None, None
