"""B200 drop-in for dcrecommend/dcue/dcue.py: Deep Content-User Embedding network
(Lee et al., DLRS 2018) with forward(u, pos, neg) on hand-written sm_100a kernels."""
import os

import torch
import torch.nn as nn

from .. import ops
from .audiomodels.truedcuemel1d import TrueDcueNetMel1D
from .audiomodels.truedcuemel1dbn import TrueDcueNetMel1DBn
from .audiomodels.truedcuemel1dres import TrueDcueNetMel1DRes
from .audiomodels.truedcuemel1dresbn import TrueDcueNetMel1DResBn
from .embeddings.userembedding import UserEmbeddings


class CosineSimilarity(nn.Module):
    """nn.CosineSimilarity(dim=1) for row-wise [M,F] x [M,F] inputs (DCUE.predict,
    dcrecommend/nn/dcue.py:513) on the score kernel."""

    def __init__(self, dim=1, eps=1e-8):
        super().__init__()
        self.dim, self.eps = dim, eps

    def forward(self, x1, x2):
        if x1.dim() != 2 or x1.shape != x2.shape or self.dim != 1:
            raise NotImplementedError("CosineSimilarity on B200 supports row-wise [M,F] x [M,F] inputs")
        M, F = x1.shape
        feats = torch.cat([x2, torch.zeros_like(x2)], dim=0)  # cos(x, 0) == 0 -> scores = cos(x1, x2)
        return ops.ScoreFn.apply(x1, feats, M, 1).view(M)


class DCUENet(nn.Module):

    """DCUE model: user tower + mel-spectrogram ConvNet song tower + cosine scores."""

    def __init__(self, dict_args):
        """dict_args keys: feature_dim, conv_hidden, user_embdim, user_count, model_type."""
        super().__init__()
        self.feature_dim = dict_args["feature_dim"]
        self.conv_hidden = dict_args["conv_hidden"]
        self.user_embdim = dict_args["user_embdim"]
        self.user_count = dict_args["user_count"]
        self.model_type = dict_args["model_type"]

        conv_args = {"output_size": self.feature_dim, "hidden_size": self.conv_hidden}
        towers = {"truedcuemel1d": TrueDcueNetMel1D, "truedcuemel1dres": TrueDcueNetMel1DRes,
                  "truedcuemel1dbn": TrueDcueNetMel1DBn, "truedcuemel1dresbn": TrueDcueNetMel1DResBn}
        if self.model_type not in towers:
            raise ValueError("{} is not a recognized model type!".format(self.model_type))
        self.conv = towers[self.model_type](conv_args)

        self.user_embd = UserEmbeddings({"user_embdim": self.user_embdim, "user_count": self.user_count,
                                         "feature_dim": self.feature_dim})
        self.sim = CosineSimilarity(dim=1)
        self._side_stream = None

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_side_stream"] = None      # CUDA streams do not pickle
        return d

    # The user tower (gather + 2-layer MLP: a dozen small latency-bound kernels) does not depend on the song tower until the
    # score kernel, so it runs on a side stream next to the tower's big kernels; autograd replays its backward on that
    # stream as well.  Inside a CUDA-graph capture this becomes a parallel branch of the graph.  DCUE_USER_STREAM=0: one stream.
    def _fork_user_tower(self, u):
        if not self.user_embd_on_side_stream(u):
            # one stream: the user tower is evaluated by _join_user_tower, i.e. AFTER the song tower.  Autograd runs the backward
            # of the later forward op first, so the user tower's backward -- and with it the data-parallel exchange of the table
            # gradient rows, which then overlaps the whole song-tower backward on a side stream (ops.UserTowerFn) -- comes first
            return None, None
        cur = torch.cuda.current_stream()
        if self._side_stream is None or self._side_stream.device != cur.device:
            self._side_stream = torch.cuda.Stream(device=cur.device)
        side = self._side_stream
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            u_f = self.user_embd(u)
        return u_f, side

    def _join_user_tower(self, u_f, side, u, feats=None):
        if u_f is None:
            # the song tower is already in the autograd graph: the user tower's backward will run before the song tower's
            # and may overlap it on a side stream (ops.UserTowerFn.backward)
            if feats is not None and feats.requires_grad and hasattr(self.user_embd, "embeddings"):
                self.user_embd._overlap_bwd = True
            return self.user_embd(u)
        if side is not None:
            cur = torch.cuda.current_stream()
            cur.wait_stream(side)
            u_f.record_stream(cur)
        return u_f

    def user_embd_on_side_stream(self, u):
        """DCUE_USER_STREAM=1 runs the user tower on a side stream next to the song tower.  Off by default: the round-2 kernel
        timeline showed its low-occupancy GEMM blocks taking register-file room on ~80 SMs, so the one-wave input pass (3
        blocks per SM by design) needed a second wave -- 554 instead of 372 us -- for 35 us of user-tower work hidden."""
        return (torch.is_tensor(u) and u.is_cuda and os.environ.get("DCUE_USER_STREAM", "0") != "0")

    def forward(self, u, pos, neg=None):
        """u int64 [B]; pos f32 [B,128,L]; neg f32 [B,N,128,L] ->
        (scores [B,N], u_featvects [B,F], pos_featvects [B,F], neg_featvects [B,N,F]).
        With neg=None the reference raises NameError; here scores is [B,1] = cos(u,pos) and
        neg_featvects is None."""
        u_featvects, side = self._fork_user_tower(u)
        B = pos.shape[0]
        if neg is not None:
            N = neg.shape[1]
            feats = self.conv.forward_posneg(pos, neg)
            u_featvects = self._join_user_tower(u_featvects, side, u, feats)
            scores = ops.ScoreFn.apply(u_featvects, feats, B, N)
            return scores, u_featvects, feats[:B], feats[B:].view(B, N, self.feature_dim)
        pos_featvects = self.conv.forward_posneg(pos, None)
        u_featvects = self._join_user_tower(u_featvects, side, u, pos_featvects)
        scores = self.sim(u_featvects, pos_featvects).view(B, 1)
        return scores, u_featvects, pos_featvects, None

    def _indexed_feats(self, pool, pos_idx, neg_idx, pos_off, neg_off, frames):
        B, N = neg_idx.shape
        idx = torch.cat([pos_idx.reshape(-1), neg_idx.reshape(-1)])
        off = None
        if pos_off is not None or neg_off is not None:
            po = torch.zeros(B, dtype=torch.int32, device=idx.device) if pos_off is None else pos_off.reshape(-1).to(torch.int32)
            no = torch.zeros(B * N, dtype=torch.int32, device=idx.device) if neg_off is None else neg_off.reshape(-1).to(torch.int32)
            off = torch.cat([po.to(idx.device), no.to(idx.device)])
        return self.conv.forward_indexed(pool, idx, off, frames)

    def forward_indexed(self, u, pool, pos_idx, neg_idx, pos_off=None, neg_off=None, frames=131):
        """forward(u, pos, neg) with pos = pool[pos_idx, :, off:off+frames], neg = pool[neg_idx, ...] taken
        from a resident device pool by index: same outputs, no dense [B,N,128,L] tensor, no H2D copy."""
        u_featvects, side = self._fork_user_tower(u)
        B, N = neg_idx.shape
        feats = self._indexed_feats(pool, pos_idx.to(pool.device), neg_idx.to(pool.device), pos_off, neg_off, frames)
        u_featvects = self._join_user_tower(u_featvects, side, u, feats)
        scores = ops.ScoreFn.apply(u_featvects, feats, B, N)
        return scores, u_featvects, feats[:B], feats[B:].view(B, N, self.feature_dim)

    def hinge_loss_step_indexed(self, u, pool, pos_idx, neg_idx, margin, pos_off=None, neg_off=None, frames=131,
                                batch_total=None):
        """hinge_loss_step on the index feed."""
        u_featvects, side = self._fork_user_tower(u)
        B, N = neg_idx.shape
        feats = self._indexed_feats(pool, pos_idx.to(pool.device), neg_idx.to(pool.device), pos_off, neg_off, frames)
        u_featvects = self._join_user_tower(u_featvects, side, u, feats)
        total = B if batch_total is None else batch_total
        loss, _ = ops.HingeLossFn.apply(u_featvects, feats, B, N, margin, total)
        return loss

    def raise_if_index_error(self):
        self.user_embd.raise_if_index_error()
        self.conv.raise_if_index_error()

    def error_flags(self):
        """The device int32 flags raised by an out-of-range user index / song index (for FusedAdam.set_skip_flags: a
        flagged step must not touch the parameters)."""
        dev = self.conv.fc.weight.device
        if self.conv._err is None or self.conv._err.device != dev:
            self.conv._err = torch.zeros(1, dtype=torch.int32, device=dev)
        flags = [self.conv._err]
        if hasattr(self.user_embd, "_err_flag"):
            flags.append(self.user_embd._err_flag())
        return flags

    def hinge_loss_step(self, u, pos, neg, margin, batch_total=None, return_all=False):
        """forward + DCUE._loss_func (max(0, margin - scores).sum(1).mean()) with the scoring, the
        loss and their backward fused in one kernel.  batch_total = global batch size under data
        parallelism (defaults to the local B)."""
        u_featvects, side = self._fork_user_tower(u)
        B, N = neg.shape[0], neg.shape[1]
        feats = self.conv.forward_posneg(pos, neg)
        u_featvects = self._join_user_tower(u_featvects, side, u, feats)
        total = B if batch_total is None else batch_total
        loss, scores = ops.HingeLossFn.apply(u_featvects, feats, B, N, margin, total)
        if return_all:
            return loss, scores, u_featvects, feats[:B], feats[B:].view(B, N, self.feature_dim)
        return loss
