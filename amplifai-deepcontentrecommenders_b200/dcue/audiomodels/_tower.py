"""Shared implementation of the four DCUE song towers on the B200 kernels.

Module tree, constructor arguments, attribute names, parameter initialisation (and therefore RNG
consumption under torch.manual_seed) and ``state_dict`` keys are those of the reference classes
(dcrecommend/dcue/audiomodels/truedcuemel1d{,bn,res,resbn}.py); the nn.Conv1d / nn.BatchNorm1d /
nn.Linear children are PARAMETER HOLDERS only — their forward is never called.  All compute goes
through ``ops.SongTowerFn`` (hand-written sm_100a kernels).
"""
import torch
from torch import nn

from ... import ops


class TowerBase(nn.Module):
    _has_bn = False
    _res = False

    def __init__(self, dict_args):
        super().__init__()
        self.output_size = dict_args["output_size"]
        self.hidden_size = dict_args["hidden_size"]
        H, F = self.hidden_size, self.output_size
        # input_size = batch size x 128 x 131
        if self._has_bn:
            self.bn0 = nn.BatchNorm1d(128)
        spec = ((128, H, 4, 2, 4, 33), (H, H, 4, 2, 4, 8), (H, H, 4, 2, 4, 2), (H, H, 2, 1, 2, 1))
        for i, (cin, cout, k, pad, pool, tlen) in enumerate(spec, start=1):
            setattr(self, "layer%d" % i, nn.Conv1d(in_channels=cin, out_channels=cout, kernel_size=k, stride=1,
                                                   padding=pad, bias=True))
            setattr(self, "relu%d" % i, nn.ReLU())
            setattr(self, "pool%d" % i, nn.MaxPool1d(kernel_size=pool))
            if self._res:
                setattr(self, "timepool%d" % i, nn.AvgPool1d(kernel_size=tlen))
            if self._has_bn:
                setattr(self, "bn%d" % i, nn.BatchNorm1d(cout))
        self.layer5 = nn.Conv1d(in_channels=H, out_channels=F, kernel_size=1, stride=1, bias=True)
        self.relu5 = nn.ReLU()
        if self._has_bn:
            self.bn5 = nn.BatchNorm1d(F)
        self.fc = nn.Linear(H * 4 + F if self._res else F, F)
        self.outsize = [H, 1]
        for i in range(1, 6):
            nn.init.kaiming_uniform_(getattr(self, "layer%d" % i).weight, nonlinearity="relu")
        nn.init.xavier_uniform_(self.fc.weight)

        names = []
        if self._has_bn:
            names += ["bn0.weight", "bn0.bias"]
        for i in range(1, 6):
            names += ["layer%d.weight" % i, "layer%d.bias" % i]
            if self._has_bn:
                names += ["bn%d.weight" % i, "bn%d.bias" % i]
        names += ["fc.weight", "fc.bias"]
        self._param_names = tuple(names)
        self._dp = None  # set by parallel.DataParallelDCUE (BatchNorm statistics all-reduce)
        self._err = None  # device flag: out-of-range song index / crop offset in the index feed

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_dp"] = None  # process-group handles do not pickle (DCUE.save pickles the trainer)
        d["_err"] = None
        return d

    def _params(self):
        return [self.get_parameter(n) for n in self._param_names]

    def forward_posneg(self, pos, neg=None):
        """Tower over the rows of `pos` [B,128,L] followed by `neg` [B,N,128,L] (or [M,128,L]),
        equal to self(torch.cat([pos, neg.view(-1,128,L)])) without the copy -> [S, F]."""
        return ops.SongTowerFn.apply(pos, neg, None, self, self.training, *self._params())

    def forward_indexed(self, pool, idx, off=None, frames=131):
        """Tower over crops of a RESIDENT song pool: spectrogram s = pool[idx[s], :, off[s]:off[s]+frames]
        (pool fp32 [n_songs,128,T] on the device, idx int64, off int32 or None = 0) -> [S, F].
        Equals self(torch.stack([pool[i, :, o:o+frames] ...])) without materialising that batch or copying
        it from the host."""
        dev = self.fc.weight.device
        if self._err is None or self._err.device != dev:
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        return ops.SongTowerFn.apply(pool, None, (idx, off, frames, self._err), self, self.training, *self._params())

    def raise_if_index_error(self):
        if self._err is not None and int(self._err.item()):
            self._err.zero_()
            raise IndexError("song index or crop offset out of range")

    def forward(self, x):
        """x [S,128,L] -> [S,F]; like the reference's trailing .squeeze(), S == 1 gives [F]."""
        out = self.forward_posneg(x, None)
        return out.squeeze()
