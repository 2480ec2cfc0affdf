"""B200 drop-in for dcrecommend/nn/dcue.py: the DCUE trainer (same constructor, attributes and
method names) driving the sm_100a kernels.

What changed relative to the reference's loops (results are the same quantities):
  * _train_epoch/_eval_epoch use DCUENet.hinge_loss_step (scores + hinge + backward fused in one
    kernel) and keep the running loss on the device: one host read per epoch instead of per step;
  * _user_factors is one batched call, _item_factors accumulates with an index_add on the device;
  * predict() gathers factor rows with tensor indexing instead of Python lists;
  * recommend_topk() is new: all users x all songs cosine scores with a fused top-k (BASELINE cfg5);
  * score()/score_song() keep every score on the device and compute AUC / mAP with one kernel over all sampled users
    (csrc/metrics.cu) instead of Python lists + sklearn per user;
  * save() stores state_dicts (loadable under torch >= 2.6), load() also accepts the reference's
    whole-object pickles; factor matrices are stored as CPU tensors like the reference's;
  * user_factors / item_factors live on the GPU while the trainer runs (the reference keeps CPU tensors): call
    .cpu() before .numpy();
  * an out-of-range user / song index does not synchronise every step: the fused optimizers skip the update of a flagged
    step on the device (parameters stay intact) and IndexError is raised when the epoch's loss is read;
  * under torchrun (torch.distributed initialised, world size > 1) _init_nn wraps the model in
    parallel.DataParallelDCUE, every rank trains on its shard of each loader (DistributedSampler), only rank 0 writes
    checkpoints; close() releases captured CUDA graphs before the process group.
"""
import os

import numpy as np
import torch
import torch.distributed as dist
from torch import optim
from torch.optim.lr_scheduler import StepLR
from torch.utils.data import DataLoader, Subset
from torch.utils.data.distributed import DistributedSampler

from .. import _lib as L
from .. import eval as dcue_eval
from ..dcue.dcue import DCUENet
from ..optim.cyclic_scheduler import CyclicLRWithRestarts
from ..optim.fused_adam import FusedAdam
from ..optim.ranger import Ranger
from .trainer import Trainer


class DCUE(Trainer):

    """Train and evaluate the DCUE model."""

    def __init__(self, feature_dim=100, conv_hidden=128, batch_size=64, neg_batch_size=20, u_embdim=300, margin=0.2,
                 optimize='adam', lr=0.00001, beta_one=0.9, beta_two=0.99, eps=1e-8, weight_decay=0, restart_period=30,
                 t_mult=2, num_epochs=90, model_type='truedcuemel1dbn', eval_pct=0.025, val_pct=1.0):
        Trainer.__init__(self)
        self.feature_dim = feature_dim
        self.conv_hidden = conv_hidden
        self.batch_size = batch_size
        self.neg_batch_size = neg_batch_size
        self.u_embdim = u_embdim
        self.margin = margin
        self.optimize = optimize
        self.lr = lr
        self.beta_one = beta_one
        self.beta_two = beta_two
        self.eps = eps
        self.weight_decay = weight_decay
        self.restart_period = restart_period
        self.t_mult = t_mult
        self.num_epochs = num_epochs
        self.model_type = model_type
        self.eval_pct = eval_pct
        self.val_pct = val_pct

        self.n_users = None
        self.n_items = None
        self.epoch_size = None

        self.model_dir = None
        self.train_data = self.val_data = self.test_data = None
        self.pred_data = self.truth_data = self.item_data = None

        self.model = None
        self.optimizer = None
        self.scheduler = None
        self.loss_func = None
        self.dict_args = None
        self.nn_epoch = 0

        self.item_factors = None
        self.user_factors = None
        self.best_item_factors = None
        self.best_user_factors = None
        self.best_val_map = 0
        self.best_val_auc = 0
        self.best_val_loss = float('inf')

        self.metadata_path = None
        self.triplets_path = None
        self.num_workers = 8

        self.USE_CUDA = torch.cuda.is_available()
        self._dp = None           # parallel.DataParallelDCUE under torchrun
        self._graph_steps = []    # captured CUDA graphs (released by close())

    # ------------------------------------------------------------------ multi-GPU plumbing
    @staticmethod
    def _world():
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    @staticmethod
    def _rank():
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    def close(self):
        """Release captured CUDA graphs (they hold NCCL work) BEFORE tearing the process group down, so a multi-GPU run
        exits through the normal path."""
        for g in self._graph_steps:
            g.release()
        self._graph_steps = []
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()

    # ------------------------------------------------------------------ model / optimiser
    def _init_nn(self, audio_model=None):
        """Build DCUENet + optimizer + SGDR scheduler (reference nn/dcue.py:129-165)."""
        self.dict_args = {'feature_dim': self.feature_dim, 'conv_hidden': self.conv_hidden, 'user_embdim': self.u_embdim,
                          'user_count': self.n_users, 'model_type': self.model_type}
        self.model = DCUENet(self.dict_args)
        if audio_model is not None:  # warm-start the song tower from a state_dict
            state_dict = self.model.state_dict()
            state_dict.update(audio_model)
            self.model.load_state_dict(state_dict)
        if not self.USE_CUDA:
            raise RuntimeError("the DCUE B200 trainer needs a CUDA device (there is no CPU fallback)")
        self.model = self.model.cuda()

        if self.optimize == 'adam':
            # torch.optim.Adam semantics and state_dict layout, one multi-tensor launch per step (optim/fused_adam.py)
            self.optimizer = FusedAdam(self.model.parameters(), self.lr, (self.beta_one, self.beta_two), self.eps,
                                       self.weight_decay)
        elif self.optimize == 'sgd':
            self.optimizer = optim.SGD(self.model.parameters(), self.lr, self.beta_one, weight_decay=self.weight_decay,
                                       nesterov=True)
            self.scheduler = StepLR(self.optimizer, 1, 1 - 1e-6)
        elif self.optimize == 'ranger':
            self.optimizer = Ranger(self.model.parameters(), lr=self.lr, alpha=0.5, k=6, N_sma_threshhold=5,
                                    betas=(self.beta_one, self.beta_two), eps=1e-5, weight_decay=self.weight_decay)
        else:
            raise ValueError("unknown optimizer {!r}".format(self.optimize))
        if hasattr(self.optimizer, 'set_skip_flags'):
            self.optimizer.set_skip_flags(self.model.error_flags())
        self._dp = None
        if self._world() > 1:
            from ..parallel import DataParallelDCUE
            self._dp = DataParallelDCUE(self.model)      # broadcasts rank 0's parameters, installs the SyncBN hooks
        # the scheduler counts GLOBAL samples per batch (every rank consumes batch_size of them)
        self.scheduler = CyclicLRWithRestarts(self.optimizer, self.batch_size * self._world(), epoch_size=self.epoch_size,
                                              restart_period=self.restart_period, t_mult=self.t_mult, policy='cosine')

    def _loss_func(self, preds):
        """max(0, margin - scores).sum(1).mean()  (reference nn/dcue.py:167-170); the training loop
        uses the fused kernel equivalent, DCUENet.hinge_loss_step."""
        return torch.max(torch.zeros_like(preds), self.margin - preds).sum(dim=1).mean()

    @staticmethod
    def _to_device(batch_samples):
        u = batch_samples['u'].cuda(non_blocking=True)
        pos = batch_samples['X'].cuda(non_blocking=True)
        neg = batch_samples['Ns'].cuda(non_blocking=True)
        return u, pos, neg

    def _train_epoch(self, loader):
        """One pass over `loader`: zero_grad, forward, hinge loss, backward, optimizer.step,
        scheduler.batch_step per batch.  Returns (samples_processed, mean train loss).  Data parallel: every rank runs
        its own batches, gradients are the global-batch gradients (parallel.DataParallelDCUE)."""
        self.model.train()
        loss_sum = torch.zeros((), device='cuda')
        samples_processed = 0
        guarded = bool(getattr(self.optimizer, 'skip_flags', None))
        for batch_samples in loader:
            u, pos, neg = self._to_device(batch_samples)
            self.model.zero_grad(set_to_none=True)
            if self._dp is not None:
                loss = self._dp.loss_step(u, pos, neg, self.margin)
                loss.backward()
                self._dp.reduce_gradients()
                loss = self._dp.reduce_loss(loss)
            else:
                loss = self.model.hinge_loss_step(u, pos, neg, self.margin)
                loss.backward()
            if not guarded:
                # optimizers without the on-device guard (SGD): a bad index must not reach the parameters
                self.model.raise_if_index_error()
            self.optimizer.step()
            self.scheduler.batch_step()
            samples_processed += pos.size()[0] * self._world()
            loss_sum += loss.detach() * (pos.size()[0] * self._world())
        self.model.raise_if_index_error()
        train_loss = loss_sum.item() / max(samples_processed, 1)
        return samples_processed, train_loss

    def _eval_epoch(self, loader):
        """Validation loss with BatchNorm running statistics."""
        self.model.eval()
        loss_sum = torch.zeros((), device='cuda')
        samples_processed = 0
        with torch.no_grad():
            for batch_samples in loader:
                u, pos, neg = self._to_device(batch_samples)
                loss = self.model.hinge_loss_step(u, pos, neg, self.margin)
                samples_processed += pos.size()[0]
                loss_sum += loss * pos.size()[0]
        self.model.user_embd.raise_if_index_error()
        return samples_processed, loss_sum.item() / max(samples_processed, 1)

    # ------------------------------------------------------------------ training driver
    def fit(self, train_dataset, val_dataset, test_dataset, pred_dataset, truth_dataset, item_dataset, n_users, n_items,
            triplets_path, metadata_path, save_dir, warm_start=False, audio_model=None):
        """Train for num_epochs sub-epochs (each = 1/10 of the training set), validating, extracting
        factors and scoring AUC / mAP after every sub-epoch, like the reference's fit()."""
        print("Settings:\n  Feature Dim: {}\n  Conv Dim: {}\n  User Embedding Dim: {}\n  Batch Size: {}\n"
              "  Negative Batch Size: {}\n  Margin: {}\n  Optimizer: {}\n  Learning Rate: {}\n  Weight Decay: {}\n"
              "  Restart Period: {}\n  T Multiplier: {}\n  Num Epochs: {}\n  Model Type: {}\n  Num Users: {}\n"
              "  Num Items: {}\n  Triplets TXT: {}\n  Metadata CSV: {}\n  Save Dir: {}".format(
                  self.feature_dim, self.conv_hidden, self.u_embdim, self.batch_size, self.neg_batch_size, self.margin,
                  self.optimize, self.lr, self.weight_decay, self.restart_period, self.t_mult, self.num_epochs,
                  self.model_type, n_users, n_items, triplets_path, metadata_path, save_dir), flush=True)
        self.epoch_size = int(int(np.ceil(len(train_dataset) / 10)) // self.batch_size) * self.batch_size
        self.n_users, self.n_items = n_users, n_items
        self.triplets_path, self.metadata_path = triplets_path, metadata_path
        self.model_dir = save_dir

        nw = self.num_workers
        truth_loader = DataLoader(truth_dataset, batch_size=1024, shuffle=True, num_workers=nw)
        val_dataset.subset(p=self.val_pct)
        val_loader = DataLoader(val_dataset, batch_size=self.batch_size, shuffle=False, num_workers=nw, pin_memory=True)
        if not warm_start:
            self._init_nn(audio_model)

        train_loss, samples_processed = 0, 0
        while self.nn_epoch < self.num_epochs + 1:
            for train_loader in self._batch_loaders(train_dataset, k=10):
                if self.nn_epoch > 0:
                    self.scheduler.step()
                    sp, train_loss = self._train_epoch(train_loader)
                    samples_processed += sp
                _, val_loss = self._eval_epoch(val_loader)
                self._user_factors(item_dataset)
                self._item_factors(item_dataset)
                pred_loader = DataLoader(pred_dataset, batch_size=1024, shuffle=True, num_workers=nw)
                val_auc, val_map = self._compute_scores('val', pred_loader, truth_loader, train_dataset, val_dataset,
                                                        test_dataset, pct=self.eval_pct)
                val_user_auc, val_user_map = self._compute_scores_song(pred_loader, pct=self.eval_pct)
                pred_loader = DataLoader(truth_dataset, batch_size=1024, shuffle=True, num_workers=nw)
                train_auc, train_map = self._compute_scores('train', pred_loader, truth_loader, train_dataset, val_dataset,
                                                            test_dataset, pct=self.eval_pct)
                print("\nEpoch: [{}/{}]\tSamples: [{}/{}]\tTrain Loss: {}\tVal Loss: {}\tTrain AUC: {}\tVal AUC: {}\t"
                      "Train mAP: {}\tVal mAP: {}\tVal UAUC: {}\tVal UmAP: {}".format(
                          self.nn_epoch, self.num_epochs, samples_processed, len(train_dataset) * self.num_epochs,
                          train_loss, val_loss, train_auc, val_auc, train_map, val_map, val_user_auc, val_user_map),
                      flush=True)
                self._update_best(val_map, val_auc, val_loss)
                self.nn_epoch += 1          # like the reference, a started sweep over the 10 loaders always finishes

    def _batch_loaders(self, dataset, k=None):
        loaders = []
        for subset_batch_indexes in dataset.get_batches(k):
            sub = Subset(dataset, subset_batch_indexes)
            if self._world() > 1:   # every rank draws a disjoint shard of the sub-epoch
                sampler = DistributedSampler(sub, self._world(), self._rank(), shuffle=True, seed=self.nn_epoch, drop_last=True)
                loaders += [DataLoader(sub, batch_size=self.batch_size, sampler=sampler, num_workers=self.num_workers,
                                       drop_last=True, pin_memory=True)]
            else:
                loaders += [DataLoader(sub, batch_size=self.batch_size, shuffle=True, num_workers=self.num_workers,
                                       drop_last=True, pin_memory=True)]
        return loaders

    # ------------------------------------------------------------------ factors
    def _user_factors(self, item_data=None, chunk=65536):
        """user_factors[U, F]: eval-mode user tower for every user index, batched on the device."""
        self.model.eval()
        dev = next(self.model.parameters()).device
        out = torch.zeros([self.n_users, self.feature_dim], device=dev)
        if item_data is not None and hasattr(item_data, 'user_index'):
            idx = torch.as_tensor(sorted(set(item_data.user_index.values())), dtype=torch.int64, device=dev)
        else:
            idx = torch.arange(self.n_users, dtype=torch.int64, device=dev)
        with torch.no_grad():
            for s in range(0, idx.numel(), chunk):
                part = idx[s:s + chunk]
                out[part] = self.model.user_embd(part)
        self.user_factors = out

    def get_item_factors(self, loader, n_iter=1):
        """Mean eval-mode tower output per song over n_iter passes (random crops) of `loader`."""
        dev = next(self.model.parameters()).device
        item_factors = torch.zeros([len(loader.dataset.songid2metaindex), self.feature_dim], device=dev)
        self.model.eval()
        with torch.no_grad():
            for _ in range(n_iter):
                for batch_samples in loader:
                    X = batch_samples['X'].to(dev, non_blocking=True)
                    idx = torch.as_tensor(batch_samples['metadata_index'], dtype=torch.int64, device=dev).view(-1)
                    f = self.model.conv.forward_posneg(X, None)
                    item_factors.index_add_(0, idx, f)
        item_factors /= n_iter
        return item_factors

    def _item_factors(self, item_data, n_iter=10):
        loader = DataLoader(item_data, batch_size=self.batch_size, shuffle=False, num_workers=min(4, self.num_workers))
        self.item_factors = self.get_item_factors(loader, n_iter=n_iter)

    def insert_best_factors(self):
        self.item_factors = self.best_item_factors
        self.user_factors = self.best_user_factors

    # ------------------------------------------------------------------ scoring
    def _pair_scores(self, loader, first_key, second_key, first_factors, second_factors):
        dev = first_factors.device
        scores, targets = [], []
        self.model.eval()
        with torch.no_grad():
            for batch_samples in loader:
                a = first_factors[torch.as_tensor(batch_samples[first_key], dtype=torch.int64, device=dev).view(-1)]
                b = second_factors[torch.as_tensor(batch_samples[second_key], dtype=torch.int64, device=dev).view(-1)]
                y = batch_samples['y']
                if b.size()[0] > 1:
                    scores += self.model.sim(a.contiguous(), b.contiguous()).cpu().numpy().tolist()
                    targets += torch.as_tensor(y).view(-1).numpy().tolist()
        return scores, targets

    def predict(self, user, loader):
        """(scores, targets) of one user against its candidate songs (cosine of factor rows)."""
        loader.dataset.create_user_data(user)
        if not loader.dataset.user_has_songs:
            return None, None
        uf, itf = self.user_factors.cuda(), self.item_factors.cuda()
        return self._pair_scores(loader, 'u', 'song_idx', uf, itf)

    def predict_song(self, song, loader):
        loader.dataset.create_song_data(song)
        if not loader.dataset.song_has_users:
            return None, None
        uf, itf = self.user_factors.cuda(), self.item_factors.cuda()
        return self._pair_scores(loader, 'u', 'song_idx', uf, itf)

    def _pair_scores_device(self, loader, first_factors, second_factors):
        """_pair_scores without leaving the device: -> (scores f32 [n], targets u8 [n]) or (None, None)."""
        dev = first_factors.device
        scores, targets = [], []
        self.model.eval()
        with torch.no_grad():
            for batch_samples in loader:
                a = first_factors[torch.as_tensor(batch_samples['u'], dtype=torch.int64, device=dev).view(-1)]
                b = second_factors[torch.as_tensor(batch_samples['song_idx'], dtype=torch.int64, device=dev).view(-1)]
                if b.size()[0] > 1:
                    scores.append(self.model.sim(a.contiguous(), b.contiguous()))
                    targets.append(torch.as_tensor(batch_samples['y']).view(-1).to(dev, non_blocking=True).ne(0).to(torch.uint8))
        if not scores:
            return torch.empty(0, device=dev), torch.empty(0, dtype=torch.uint8, device=dev)
        return torch.cat(scores), torch.cat(targets)

    @staticmethod
    def ranking_metrics(scores, targets, seg_offsets, group=None):
        """AUC / AP of every segment on the device (csrc/metrics.cu) -> double [n_seg, 8], see dcue_auc_ap_segments."""
        n_seg = seg_offsets.numel() - 1
        out = torch.zeros(n_seg, 8, dtype=torch.float64, device=scores.device)
        if n_seg > 0:
            L.call("dcue_auc_ap_segments", scores.contiguous().data_ptr(), targets.contiguous().data_ptr(),
                   L.ptr(None if group is None else group.contiguous()), seg_offsets.contiguous().data_ptr(), n_seg,
                   out.data_ptr(), L.stream())
        return out

    def score(self, users, pred_loader, truth_loader, k=10000):
        """Mean weighted AUC and mAP over `users`, mixing each user's positives of one split with the
        negatives of the other (the reference's estimator, nn/dcue.py:380-449).  Every user's candidate scores stay on
        the device; one kernel computes all users' AUCs (per half) and APs."""
        uf, itf = self.user_factors.cuda(), self.item_factors.cuda()
        sc, tg, gr, offs = [], [], [], [0]
        for user_id in users:
            pred_loader.dataset.create_user_data(user_id)
            if not pred_loader.dataset.user_has_songs:
                break                                   # like the reference: predict() returned (None, None)
            sp, tp = self._pair_scores_device(pred_loader, uf, itf)
            truth_loader.dataset.create_user_data(user_id)
            if truth_loader.dataset.user_has_songs:
                st, tt = self._pair_scores_device(truth_loader, uf, itf)
            else:
                st, tt = sp[:0], tp[:0]
            # half 0 = pred positives + truth negatives, half 1 = pred negatives + truth positives (nn/dcue.py:405-418)
            sc += [sp, st]
            tg += [tp, tt]
            gr += [1 - tp, tt]
            offs.append(offs[-1] + sp.numel() + st.numel())
        if len(offs) == 1:
            return np.nan, np.nan
        seg = torch.tensor(offs, dtype=torch.int64, device=uf.device)
        m = self.ranking_metrics(torch.cat(sc), torch.cat(tg), seg, torch.cat(gr))
        total = (m[:, 2] + m[:, 3]).clamp_min(1.0)
        auc = (m[:, 2] / total) * m[:, 0] + (m[:, 3] / total) * m[:, 1]
        return auc.mean().item(), m[:, 6].mean().item()

    def score_song(self, songs, pred_loader, k=10000):
        """Mean AUC / mAP over `songs` (each against its candidate users), nn/dcue.py:451-476, on the device."""
        uf, itf = self.user_factors.cuda(), self.item_factors.cuda()
        sc, tg, offs = [], [], [0]
        for song_id in songs:
            pred_loader.dataset.create_song_data(song_id)
            if not pred_loader.dataset.song_has_users:
                continue
            s, t = self._pair_scores_device(pred_loader, uf, itf)
            sc.append(s)
            tg.append(t)
            offs.append(offs[-1] + s.numel())
        if len(offs) == 1:
            return np.nan, np.nan
        seg = torch.tensor(offs, dtype=torch.int64, device=uf.device)
        m = self.ranking_metrics(torch.cat(sc), torch.cat(tg), seg)
        n, P = m[:, 2], m[:, 4]
        ap = torch.where(P == n, torch.ones_like(n), torch.where(P == 0, torch.zeros_like(n), m[:, 6]))
        return m[:, 0].mean().item(), ap.mean().item()

    def _compute_scores(self, split, pred_loader, truth_loader, train_data, val_data, test_data, pct=0.025):
        if split == 'train':
            users = list(train_data.uniq_users)
        elif split == 'val':
            users = list(set(train_data.uniq_users).intersection(set(val_data.uniq_users)))
        elif split == 'test':
            users = list(set(train_data.uniq_users).intersection(set(test_data.uniq_users)))
        else:
            raise ValueError(split)
        sample = np.random.choice(users, int(len(users) * pct)) if pct < 1 else users
        return self.score(sample, pred_loader, truth_loader)

    def _compute_scores_song(self, pred_loader, pct=0.025):
        songs = list(pred_loader.dataset.uniq_songs)
        sample = np.random.choice(songs, int(len(songs) * pct)) if pct < 1 else songs
        return self.score_song(sample, pred_loader)

    def recommend_topk(self, k=100, users=None):
        """Top-k songs for every user (or the given user indices): all-pairs cosine scores of the
        factor matrices with the fused score-GEMM + top-k kernel.  -> (scores [U,k], song index [U,k])."""
        uf = self.user_factors.cuda()
        if users is not None:
            uf = uf[torch.as_tensor(users, dtype=torch.int64, device=uf.device)]
        return dcue_eval.topk_scores(uf.contiguous(), self.item_factors.cuda().contiguous(), k)

    # ------------------------------------------------------------------ checkpoints
    def _update_best(self, val_map, val_auc, val_loss):
        if val_map > self.best_val_map:
            self.best_val_map, self.best_val_auc, self.best_val_loss = val_map, val_auc, val_loss
            self.best_item_factors = self.item_factors.clone()
            self.best_user_factors = self.user_factors.clone()
            self.save(models_dir=self.model_dir)
        elif self.nn_epoch % 5 == 0:
            self.save(models_dir=self.model_dir)

    def _format_model_subdir(self):
        return "DCUE_fd_{}_ch_{}_uh_{}_op_{}_lr_{}_wd_{}_rp_{}_tm_{}_nu_{}_ni_{}_mt_{}".format(
            self.feature_dim, self.conv_hidden, self.u_embdim, self.optimize, self.lr, self.weight_decay,
            self.restart_period, self.t_mult, self.n_users, self.n_items, self.model_type)

    _STATE_KEYS = ('model', 'optimizer', 'scheduler', 'train_data', 'val_data', 'test_data', 'pred_data', 'truth_data',
                   'item_data', 'loss_func')

    def save(self, models_dir=None):
        """<models_dir>/<subdir>/epoch_<n>.pth with hyper-parameters, factor matrices and the
        state_dicts of model / optimizer / scheduler."""
        if self.model is None or models_dir is None or self._rank() != 0:
            return
        path = os.path.join(models_dir, self._format_model_subdir())
        os.makedirs(path, exist_ok=True)
        ckpt = {k: v for k, v in self.__dict__.items() if k not in self._STATE_KEYS and not k.startswith('_')}
        for k in ('user_factors', 'item_factors', 'best_user_factors', 'best_item_factors'):
            if torch.is_tensor(ckpt.get(k)):
                ckpt[k] = ckpt[k].cpu()                      # the reference stores CPU factor matrices
        ckpt['format'] = 'dcue_b200.state_dict.v1'
        ckpt['model_state'] = self.model.state_dict()
        ckpt['optimizer_state'] = self.optimizer.state_dict()
        ckpt['scheduler_state'] = self.scheduler.state_dict()
        torch.save(ckpt, os.path.join(path, "epoch_{}.pth".format(self.nn_epoch)))

    def load(self, model_dir, epoch):
        """Restore a checkpoint written by save() (or by the reference's save(), which pickles the
        whole trainer) and rebuild model / optimizer / scheduler; nn_epoch advances by one."""
        model_file = os.path.join(model_dir, "epoch_{}.pth".format(epoch))
        checkpoint = torch.load(model_file, map_location='cuda' if torch.cuda.is_available() else 'cpu',
                                weights_only=False)
        states = {}
        if checkpoint.get('format') == 'dcue_b200.state_dict.v1':
            for k in ('model', 'optimizer', 'scheduler'):
                states[k] = checkpoint.pop(k + '_state')
            checkpoint.pop('format')
        else:
            for k in ('model', 'optimizer', 'scheduler'):
                states[k] = checkpoint.pop(k).state_dict()
        for k, v in checkpoint.items():
            setattr(self, k, v)
        self.USE_CUDA = torch.cuda.is_available()
        self._init_nn()
        self.model.load_state_dict(states['model'])
        self.optimizer.load_state_dict(states['optimizer'])
        self.scheduler.load_state_dict(states['scheduler'])
        self.nn_epoch += 1
