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

import os

import torch
import torch.distributed as dist


def shard_slice(n, rank, world):
    """Contiguous slice [lo, hi) of n items owned by `rank` (rank r gets rows
    [r*B/W, (r+1)*B/W))."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def row_exchange():
    """Replicated user table under DP: exchange gradient rows (default) or all-reduce the dense gradient
    (DCUE_DP_DENSE_TABLE=1, the A/B switch)."""
    return os.environ.get("DCUE_DP_DENSE_TABLE", "0") != "1"


def flat_bucket_names(named_grads):
    """Names of the gradients that need the SUM all-reduce: everything except the affine parameters
    of bn1..bn5, whose gradients are already global (computed from all-reduced sums).  bn0 is folded
    into layer1, so its gradients are local partial sums like any weight gradient.  A row-sharded user
    table's gradient is complete on its owner and is excluded as well; so is the replicated table's dense gradient,
    which UserTowerFn builds from the all-gathered gradient rows of every rank (already the global sum)."""
    return [n for n, _ in named_grads if (".bn" not in n or ".bn0." in n) and not n.endswith("user_embd.shard")
            and not (n.endswith("user_embd.embeddings.weight") and row_exchange())]


class PeerAllReduce:
    """Tiny fp64 all-reduce over NVLink peer memory (csrc/peer.cu) on torch's symmetric-memory buffers."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self.L = L
        self.slot = L.lib().dcue_peer_allreduce_slot_doubles()
        pg = group if group is not None else dist.group.WORLD
        self.buf = symm.empty(2 * self.slot, dtype=torch.float64, device=device)
        self.hdl = symm.rendezvous(self.buf, pg.group_name)
        self.counter = torch.zeros(1, dtype=torch.int32, device=device)
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        if self.hdl.signal_pad_size < 4 * self.world:
            raise RuntimeError("signal pad too small")

    def __call__(self, t):
        self.L.call("dcue_peer_allreduce_f64", self.hdl.buffer_ptrs_dev, self.hdl.signal_pad_ptrs_dev, self.counter.data_ptr(),
                    self.rank, self.world, t.data_ptr(), t.numel(), self.L.stream())


class DataParallelDCUE:
    """Wraps a DCUENet for data-parallel training on the current process group."""

    def __init__(self, model, group=None, broadcast=True):
        self.model, self.group = model, group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        model.conv._dp = self if self.world_size > 1 else None
        # SyncBN statistics: 13 all-reduces of 2 KB per step -> one-shot peer-memory kernel instead of NCCL
        # (DCUE_DP_PEER_ALLREDUCE=0, a CPU/gloo group or missing peer access keep the NCCL path)
        self._peer = None
        params = list(model.parameters())
        if (self.world_size > 1 and params and params[0].is_cuda and dist.get_backend(group) == "nccl"
                and os.environ.get("DCUE_DP_PEER_ALLREDUCE", "1") != "0"):
            try:
                self._peer = PeerAllReduce(group, params[0].device)
            except Exception as exc:  # noqa: BLE001  (no peer access / symmetric memory unavailable)
                print("DataParallelDCUE: peer all-reduce unavailable (%s); using NCCL for the BatchNorm statistics" % exc)
        if hasattr(model.user_embd, "embeddings"):      # replicated table: row exchange instead of a dense all-reduce
            model.user_embd._dp = self if (self.world_size > 1 and row_exchange()) else None
        if broadcast and self.world_size > 1:
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, 0, group=group)

    # hook used by ops.SongTowerFn for BatchNorm statistics
    def all_reduce_sum(self, t):
        if self.world_size > 1:
            if self._peer is not None and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self._peer.slot:
                self._peer(t)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def all_gather_rows(self, t):
        """[n, ...] on every rank -> [world*n, ...] in rank order."""
        t = t.contiguous()
        out = torch.empty((self.world_size * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out

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


# ------------------------------------------------------------------------------ row-sharded user table (cfg4)
def shard_rows(n_rows, rank, world):
    """Contiguous block of table rows owned by `rank`: [lo, hi)."""
    return shard_slice(n_rows, rank, world)


def local_row_index(all_idx, lo, hi):
    """Map global row indices to this owner's local rows; rows owned by other ranks map to the sentinel
    (hi - lo).  Returns (local_idx int64, owned bool)."""
    owned = (all_idx >= lo) & (all_idx < hi)
    local = torch.where(owned, all_idx - lo, torch.full_like(all_idx, hi - lo))
    return local, owned


class _ShardedRowsFn(torch.autograd.Function):
    """rows[b] = table[u[b]] for a table whose rows are block-sharded over the ranks.
    forward : all_gather(indices) -> every owner gathers the rows it holds (zeros elsewhere)
              -> all_to_all of the gathered rows back to the requesting ranks -> sum over owners
    backward: all_gather(indices, gradient rows) -> every owner segment-sums the rows it owns into its
              dense shard gradient (already the GLOBAL sum: it must not be all-reduced again)."""

    @staticmethod
    def forward(ctx, u, shard, lo, hi, group):
        from . import ops
        world = dist.get_world_size(group)
        B, E = u.numel(), shard.shape[1]
        all_idx = torch.empty(world * B, dtype=torch.int64, device=u.device)
        dist.all_gather_into_tensor(all_idx, u.contiguous().view(-1), group=group)
        local, owned = local_row_index(all_idx, lo, hi)
        safe = torch.where(owned, local, torch.zeros_like(local))
        part = ops.gather_rows(shard, safe) * owned.unsqueeze(1).to(shard.dtype)      # [world*B, E]
        recv = torch.empty(world, B, E, dtype=shard.dtype, device=u.device)
        dist.all_to_all_single(recv.view(world * B, E), part, group=group)           # row exchange over NVLink
        ctx.save_for_backward(all_idx)
        ctx.meta = (lo, hi, group, B, E)
        return recv.sum(dim=0)                                                        # exactly one owner contributes

    @staticmethod
    def backward(ctx, grows):
        from . import ops
        (all_idx,) = ctx.saved_tensors
        lo, hi, group, B, E = ctx.meta
        world = dist.get_world_size(group)
        all_g = torch.empty(world * B, E, dtype=grows.dtype, device=grows.device)
        dist.all_gather_into_tensor(all_g, grows.contiguous(), group=group)
        local, _ = local_row_index(all_idx, lo, hi)
        gshard = ops.scatter_rows(local, all_g, hi - lo)
        return None, gshard, None, None, None


class ShardedUserTable(torch.nn.Module):
    """Row-sharded replacement for UserEmbeddings' lookup table (BASELINE cfg4: 1M users x 300 over 8 GPUs).
    Each rank owns rows [lo, hi) (+ its optimizer state); the MLP stays replicated / data parallel."""

    def __init__(self, user_embd, group=None):
        super().__init__()
        from . import ops
        self._ops = ops
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        full = user_embd.embeddings.weight.detach()
        self.user_count = full.shape[0]
        self.lo, self.hi = shard_rows(self.user_count, self.rank, self.world_size)
        self.shard = torch.nn.Parameter(full[self.lo:self.hi].clone())
        self.linear1, self.linear2 = user_embd.linear1, user_embd.linear2

    def forward(self, user_idx):
        shape = user_idx.shape
        rows = _ShardedRowsFn.apply(user_idx.reshape(-1).to(self.shard.device), self.shard, self.lo, self.hi, self.group)
        out = self._ops.UserMLPFn.apply(rows, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
        return out.view(*shape, -1)

    def raise_if_index_error(self):
        pass

    def gather_full_table(self):
        """[U,E] table on every rank (for checkpoints in the reference's state_dict layout)."""
        sizes = [shard_rows(self.user_count, r, self.world_size) for r in range(self.world_size)]
        m = max(h - l for l, h in sizes)
        pad = torch.zeros(m, self.shard.shape[1], dtype=self.shard.dtype, device=self.shard.device)
        pad[: self.hi - self.lo] = self.shard.detach()
        out = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(out, pad, group=self.group)
        return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)])


def shard_user_table(model, group=None):
    """Replace model.user_embd by its row-sharded version (call before building the optimizer)."""
    model.user_embd = ShardedUserTable(model.user_embd, group)
    return model


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
