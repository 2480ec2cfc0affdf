"""Multi-GPU layer: one process per GPU, torch.distributed (NCCL over NVLink) for the plumbing.

Training is data parallel (BASELINE cfg3): each rank runs the kernels on its slice of the global
batch.  To equal the single-process reference at the GLOBAL batch size,
  * BatchNorm batch statistics and the two BatchNorm-backward sums are all-reduced per layer
    (6 layers x 2x128 fp64 each way -- SyncBN semantics; done inside ops.SongTowerFn through the
    ``all_reduce_sum`` hook installed here),
  * the hinge loss divides by the global batch (``hinge_loss_step(batch_total=...)``), so local
    gradients are partial sums and one flat SUM all-reduce of all non-BatchNorm gradients (tower,
    user MLP, the dense table gradient, and bn0 which is folded into layer1) yields exactly the reference's gradient on every rank.
Eval (BASELINE cfg5) shards songs across ranks and merges per-rank top-k lists.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_slice(n, rank, world):
    """Contiguous slice [lo, hi) of n items owned by `rank` (rank r gets rows
    [r*B/W, (r+1)*B/W))."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flat_bucket_names(named_grads):
    """Names of the gradients that need the SUM all-reduce: everything except the affine parameters
    of bn1..bn5, whose gradients are already global (computed from all-reduced sums).  bn0 is folded
    into layer1, so its gradients are local partial sums like any weight gradient."""
    return [n for n, _ in named_grads if ".bn" not in n or ".bn0." in n]


class DataParallelDCUE:
    """Wraps a DCUENet for data-parallel training on the current process group."""

    def __init__(self, model, group=None, broadcast=True):
        self.model, self.group = model, group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        model.conv._dp = self if self.world_size > 1 else None
        if broadcast and self.world_size > 1:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, 0, group=group)

    # hook used by ops.SongTowerFn for BatchNorm statistics
    def all_reduce_sum(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def loss_step(self, u, pos, neg, margin):
        """Local slice of the global batch -> loss contribution whose gradients sum to the global
        gradient.  Returns the local partial loss (sum over ranks == reference loss)."""
        return self.model.hinge_loss_step(u, pos, neg, margin, batch_total=pos.shape[0] * self.world_size)

    def loss_step_indexed(self, u, pool, pos_idx, neg_idx, margin, pos_off=None, neg_off=None, frames=131):
        """loss_step on the index feed (resident song pool)."""
        return self.model.hinge_loss_step_indexed(u, pool, pos_idx, neg_idx, margin, pos_off, neg_off, frames,
                                                  batch_total=pos_idx.shape[0] * self.world_size)

    def reduce_gradients(self):
        """One flat SUM all-reduce over all non-BatchNorm gradients."""
        if self.world_size == 1:
            return
        named = [(n, p) for n, p in self.model.named_parameters() if p.grad is not None]
        keep = set(flat_bucket_names(named))
        grads = [p.grad for n, p in named if n in keep]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in grads]), grads)])

    def reduce_loss(self, loss):
        if self.world_size > 1:
            loss = loss.detach().clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss


def sharded_topk(user_factors, item_factors_local, k, item_offset, group=None):
    """Song-sharded eval: every rank scores all users against ITS songs, then the per-rank top-k
    lists are all-gathered and merged (k-way merge kernel).  Returns the global top-k on every rank."""
    from . import eval as ev
    s, i = ev.topk_scores(user_factors, item_factors_local, k, item_offset=item_offset)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return s, i
    ss = [torch.empty_like(s) for _ in range(world)]
    ii = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(ss, s, group=group)
    dist.all_gather(ii, i, group=group)
    return ev.merge_topk(ss, ii)
