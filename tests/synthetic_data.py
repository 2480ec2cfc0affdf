"""Tiny in-memory datasets implementing the protocol the reference trainer expects from
DCUEDataset / DCUEPredset / DCUEItemset (dict samples with keys u, y, X, Ns / u, y, song_idx /
X, metadata_index).  The real datasets need the MSD audio corpus, which is out of scope."""
import numpy as np
import torch
from torch.utils.data import Dataset


class SynthWorld:
    def __init__(self, n_users=24, n_songs=40, negs=3, frames=131, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.n_users, self.n_songs, self.negs = n_users, n_songs, negs
        self.mels = torch.randn(n_songs, 128, frames, generator=g)
        rng = np.random.RandomState(seed)
        self.likes = {u: set(rng.choice(n_songs, 6, replace=False).tolist()) for u in range(n_users)}
        self.pairs = [(u, s) for u in range(n_users) for s in sorted(self.likes[u])]
        self.rng = rng


class SynthTrainSet(Dataset):
    def __init__(self, world, pairs=None):
        self.w = world
        self.pairs = list(world.pairs if pairs is None else pairs)
        self.uniq_users = sorted({u for u, _ in self.pairs})
        self.uniq_songs = sorted({s for _, s in self.pairs})

    def __len__(self):
        return len(self.pairs)

    def subset(self, p=1.0):
        pass

    def get_batches(self, k):
        idx = np.arange(len(self.pairs))
        return [list(c) for c in np.array_split(idx, k) if len(c)]

    def __getitem__(self, i):
        u, s = self.pairs[i]
        non = [x for x in range(self.w.n_songs) if x not in self.w.likes[u]]
        rng = np.random.RandomState(1000 + i)
        ns = rng.choice(non, self.w.negs)
        return {'u': torch.tensor(u), 'y': -torch.ones(self.w.negs), 'X': self.w.mels[s], 'Ns': self.w.mels[ns]}


class SynthPredSet(Dataset):
    """Per-user / per-song candidate lists: positives of this split + every non-interacted song."""

    def __init__(self, world, pairs=None):
        self.w = world
        self.pairs = list(world.pairs if pairs is None else pairs)
        self.uniq_users = sorted({u for u, _ in self.pairs})
        self.uniq_songs = sorted({s for _, s in self.pairs})
        self.rows = []
        self.user_has_songs = False
        self.song_has_users = False
        self.create_user_data(self.uniq_users[0])   # like DCUEPredset: never empty

    def create_user_data(self, user):
        pos = [s for u, s in self.pairs if u == user]
        neg = [s for s in range(self.w.n_songs) if s not in self.w.likes[user]]
        self.rows = [(user, s, 1) for s in pos] + [(user, s, 0) for s in neg]
        self.user_has_songs = len(pos) > 0

    def create_song_data(self, song):
        pos = [u for u, s in self.pairs if s == song]
        neg = [u for u in range(self.w.n_users) if song not in self.w.likes[u]]
        self.rows = [(u, song, 1) for u in pos] + [(u, song, 0) for u in neg]
        self.song_has_users = len(pos) > 0

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        u, s, y = self.rows[i]
        return {'u': torch.tensor(u), 'song_idx': torch.tensor(s), 'y': torch.tensor(y)}


class SynthItemSet(Dataset):
    def __init__(self, world):
        self.w = world
        self.user_index = {"user%d" % u: u for u in range(world.n_users)}
        self.songid2metaindex = {"song%d" % s: s for s in range(world.n_songs)}

    def __len__(self):
        return self.w.n_songs

    def __getitem__(self, i):
        return {'X': self.w.mels[i], 'metadata_index': i}
