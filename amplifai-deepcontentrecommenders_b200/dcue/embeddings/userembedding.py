"""B200 drop-in for dcrecommend/dcue/embeddings/userembedding.py."""
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

    def forward(self, user_idx):
        """user_idx: int64 tensor of any shape -> [..., feature_dim]."""
        return ops.UserTowerFn.apply(user_idx, self.embeddings.weight, self.linear1.weight, self.linear1.bias,
                                     self.linear2.weight, self.linear2.bias)
