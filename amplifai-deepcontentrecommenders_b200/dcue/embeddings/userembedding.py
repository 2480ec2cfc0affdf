"""B200 drop-in for dcrecommend/dcue/embeddings/userembedding.py."""
import torch
import torch.nn as nn

from ... import ops


class UserEmbeddings(nn.Module):

    """Embed users as feature vectors: table gather -> ReLU -> Linear -> ReLU -> Linear."""

    def __init__(self, dict_args):
        super().__init__()
        self.user_embdim = dict_args["user_embdim"]
        self.user_count = dict_args["user_count"]
        self.feature_dim = dict_args["feature_dim"]
        # parameter holders (same construction order / init as the reference, :27-31)
        self.embeddings = nn.Embedding(self.user_count, self.user_embdim)
        self.relu1 = nn.ReLU()
        self.linear1 = nn.Linear(self.user_embdim, self.user_embdim)
        self.relu2 = nn.ReLU()
        self.linear2 = nn.Linear(self.user_embdim, self.feature_dim)

        self._err = None  # device flag set by the gather kernel on an out-of-range index
        self._dp = None   # set by parallel.DataParallelDCUE: exchange gradient rows instead of the dense table gradient

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_err"] = None
        d["_dp"] = None
        return d

    def _err_flag(self):
        dev = self.embeddings.weight.device
        if self._err is None or self._err.device != dev:
            self._err = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._err

    def forward(self, user_idx):
        """user_idx: int64 tensor of any shape -> [..., feature_dim].  An out-of-range index (where
        nn.Embedding raises IndexError) yields NaN rows and sets a device flag without a host sync;
        call raise_if_index_error() -- the trainer does, whenever it reads the loss."""
        overlap, self._overlap_bwd = getattr(self, "_overlap_bwd", False), False     # set by DCUENet for this one call
        return ops.UserTowerFn.apply(user_idx, self.embeddings.weight, self.linear1.weight, self.linear1.bias,
                                     self.linear2.weight, self.linear2.bias, self._err_flag(), getattr(self, "_dp", None), overlap)

    def raise_if_index_error(self):
        if self._err is not None and int(self._err.item()):
            self._err.zero_()
            raise IndexError("index out of range in self")
