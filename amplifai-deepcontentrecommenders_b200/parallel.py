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
        self.buf = symm.empty(int(L.lib().dcue_peer_allreduce_buffer_doubles()), dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, pg.group_name)
        self.counter = torch.zeros(2, dtype=torch.int32, device=device)     # [call counter, time-out flag]
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        if self.hdl.signal_pad_size < 4 * self.world:
            raise RuntimeError("signal pad too small")

    def __call__(self, t):
        self.L.call("dcue_peer_allreduce_f64", self.hdl.buffer_ptrs_dev, self.hdl.signal_pad_ptrs_dev, self.counter.data_ptr(),
                    self.rank, self.world, t.data_ptr(), t.numel(), self.L.stream())

    def check(self):
        """Raise if a peer did not answer within the kernel's time-out (a rank died or raised): host sync."""
        if int(self.counter[1].item()):
            raise RuntimeError("peer all-reduce timed out waiting for another rank")


class PeerExchange:
    """Symmetric exchange buffer for (user index, gradient row) pairs over NVLink peer memory (csrc/peer.cu): every rank
    writes its rows into its own slot, one single-CTA kernel publishes the indices + barriers + collects all index lists,
    and the owner reads the rows it needs straight from the peers' slots."""

    def __init__(self, group, device, capacity, E):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self.L = L
        pg = group if group is not None else dist.group.WORLD
        capacity = (capacity + 1) // 2 * 2
        self.capacity, self.E = capacity, E
        nbytes = L.query("dcue_peer_exchange_bytes", capacity, E)
        self.buf = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, pg.group_name)
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        off = L.query("dcue_peer_exchange_rows_offset", capacity)
        self.rows = self.buf[off: off + capacity * E * 4].view(torch.float32).view(capacity, E)   # this rank's rows slot
        self.counter = torch.zeros(3, dtype=torch.int32, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)          # every rank's flags are zero before the first exchange

    def barrier(self, channel=0):
        self.L.call("dcue_peer_exchange_i64", self.hdl.buffer_ptrs_dev, self.counter.data_ptr(), channel, self.rank, self.world,
                    None, 0, None, self.L.stream())

    def exchange_indices(self, idx, channel=1):
        """Publish idx [B] int64, barrier, -> all ranks' indices [world*B] in rank order."""
        idx = idx.contiguous()
        B = idx.numel()
        if B > self.capacity:
            raise ValueError("exchange capacity %d < batch %d" % (self.capacity, B))
        out = torch.empty(self.world * B, dtype=torch.int64, device=idx.device)
        self.L.call("dcue_peer_exchange_i64", self.hdl.buffer_ptrs_dev, self.counter.data_ptr(), channel, self.rank, self.world,
                    idx.data_ptr(), B, out.data_ptr(), self.L.stream())
        return out

    def scatter_add(self, all_idx, B, lo, hi):
        """Dense [hi-lo, E] sum of every rank's rows slot entries whose index falls in [lo, hi)."""
        out = torch.zeros(hi - lo, self.E, dtype=torch.float32, device=all_idx.device)
        ws = torch.empty(2 * (hi - lo), dtype=torch.int32, device=all_idx.device)
        self.L.call("dcue_peer_scatter_add_rows", self.hdl.buffer_ptrs_dev, self.capacity, all_idx.data_ptr(), self.world, B, lo, hi,
                    self.E, out.data_ptr(), ws.data_ptr(), ws.numel() * 4, self.L.stream())
        return out

    def check(self):
        if int(self.counter[2].item()):
            raise RuntimeError("peer exchange timed out waiting for another rank")


class PeerGradReduce:
    """Flat all-reduce of the step's gradient tensors over NVLink peer memory (csrc/peer.cu dcue_peer_allreduce_grads)."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm
        from . import _lib as L
        self.L = L
        pg = group if group is not None else dist.group.WORLD
        self.buf = symm.empty(L.query("dcue_peer_grads_bytes"), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, pg.group_name)
        self.rank, self.world = self.hdl.rank, self.hdl.world_size
        self.counter = torch.zeros(3, dtype=torch.int32, device=device)
        self.max_elems = int(L.lib().dcue_peer_grads_max_elems())
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def __call__(self, grads):
        rows = []
        for g in grads:
            if g.dtype != torch.float32 or not g.is_contiguous():
                raise RuntimeError("PeerGradReduce needs contiguous fp32 gradients")
            rows += [g.data_ptr(), g.numel()]
        table = torch.tensor(rows, dtype=torch.int64)          # host: the list is passed to the kernel by value
        self.L.call("dcue_peer_allreduce_grads", self.hdl.buffer_ptrs_dev, self.counter.data_ptr(), self.rank, self.world,
                    table.data_ptr(), len(grads), self.L.stream())

    def check(self):
        if int(self.counter[2].item()):
            raise RuntimeError("peer gradient all-reduce timed out waiting for another rank")


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
        self._xch = None          # PeerExchange for the table-gradient rows, created on first use (needs the batch size)
        self._xstream, self._xpending = None, None      # side stream of the row exchange and its pending join
        self._gred = None         # PeerGradReduce for the flat gradient bucket
        # one-shot peer all-reduce of the flat gradient bucket: +1.1 % step rate on 2 GPUs, equal to NCCL's LL ring + cat/copy
        # on 8 (2.770 vs 2.780 ms per step, round 2); DCUE_DP_PEER_GRADS=0 selects NCCL
        if self._peer is not None and os.environ.get("DCUE_DP_PEER_GRADS", "1") != "0":
            try:
                self._gred = PeerGradReduce(group, params[0].device)
            except Exception as exc:  # noqa: BLE001
                print("DataParallelDCUE: peer gradient all-reduce unavailable (%s); using NCCL" % exc)
        if hasattr(model.user_embd, "embeddings"):      # replicated table: row exchange instead of a dense all-reduce
            model.user_embd._dp = self if (self.world_size > 1 and row_exchange()) else None
        if broadcast and self.world_size > 1:
            # replicated state only: a row-sharded table's shard is rank-specific by construction
            for n, t in list(model.named_parameters()) + list(model.named_buffers()):
                if not n.endswith("user_embd.shard"):
                    dist.broadcast(t.data, 0, group=group)

    # hook used by ops.SongTowerFn for BatchNorm statistics
    def all_reduce_sum(self, t):
        if self.world_size > 1:
            if self._peer is not None and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self._peer.slot:
                self._peer(t)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def table_exchange(self, B, E, device):
        """PeerExchange for B (index, gradient row) pairs per rank, or None when peer memory is not in use (then the rows
        travel by NCCL all-gather)."""
        if self._peer is None or os.environ.get("DCUE_DP_PEER_ROWS", "1") == "0":
            return None
        if self._xch is None or self._xch.capacity < B or self._xch.E != E:
            self._xch = PeerExchange(self.group, device, B, E)
        return self._xch

    def exchange_stream(self, table):
        """Side stream for the table-gradient row exchange of this backward pass, or None (exchange on the current stream):
        DCUE_DP_EXCHANGE_STREAM=0, or autograd is going to ACCUMULATE into an existing table.grad (see ops.UserTowerFn)."""
        if os.environ.get("DCUE_DP_EXCHANGE_STREAM", "1") == "0" or table.grad is not None:
            return None
        if self._xstream is None:
            self._xstream = torch.cuda.Stream(device=table.device)
        return self._xstream

    def exchange_pending(self, stream, keep):
        self._xpending = (stream, keep)      # tensors the side stream still reads stay referenced until the join

    def join_exchange(self):
        """The current stream waits for the side-stream row exchange of the last backward (no-op when none is pending)."""
        if self._xpending is not None:
            torch.cuda.current_stream().wait_stream(self._xpending[0])
            self._xpending = None

    def check_peers(self):
        """Host-side check of the peer kernels' time-out flags (a rank that died would otherwise go unnoticed)."""
        if self._peer is not None:
            self._peer.check()
        if self._xch is not None:
            self._xch.check()
        if self._gred is not None:
            self._gred.check()

    def all_gather_rows(self, t):
        """[n, ...] on every rank -> [world*n, ...] in rank order."""
        t = t.contiguous()
        out = torch.empty((self.world_size * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    def loss_step(self, u, pos, neg, margin):
        """Local slice of the global batch -> loss contribution whose gradients sum to the global
        gradient.  Returns the local partial loss (sum over ranks == reference loss)."""
        self.join_exchange()            # a backward whose reduce_gradients() was skipped must not leave the stream dangling
        return self.model.hinge_loss_step(u, pos, neg, margin, batch_total=pos.shape[0] * self.world_size)

    def loss_step_indexed(self, u, pool, pos_idx, neg_idx, margin, pos_off=None, neg_off=None, frames=131):
        """loss_step on the index feed (resident song pool)."""
        self.join_exchange()
        return self.model.hinge_loss_step_indexed(u, pool, pos_idx, neg_idx, margin, pos_off, neg_off, frames,
                                                  batch_total=pos_idx.shape[0] * self.world_size)

    def reduce_gradients(self):
        """One flat SUM all-reduce over all non-BatchNorm gradients (+ the join of the side-stream table-row exchange)."""
        if self.world_size == 1:
            return
        from . import ops
        ops.join_backward_side()            # user-tower gradients computed on the backward side stream are in the flat bucket
        try:
            self._reduce_flat()
        finally:
            self.join_exchange()        # after the flat bucket has been launched: the two exchanges overlap as well

    def _reduce_flat(self):
        named = [(n, p) for n, p in self.model.named_parameters() if p.grad is not None]
        keep = set(flat_bucket_names(named))
        grads = [p.grad for n, p in named if n in keep]
        if (self._gred is not None and len(grads) <= 63 and sum(g.numel() for g in grads) <= self._gred.max_elems
                and all(g.dtype == torch.float32 and g.is_contiguous() for g in grads)):
            self._gred(grads)       # one kernel over NVLink peer memory; sums land in the gradient tensors
            return
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


def owner_of_rows(idx, n_rows, world):
    """Owner rank of every global row index under shard_slice's block partition (the arithmetic of csrc/peer.cu
    block_owner) and the row's position inside the owner's shard."""
    base, rem = divmod(n_rows, world)
    cut = (base + 1) * rem
    small = idx < cut
    own_a = torch.div(idx, base + 1, rounding_mode="floor")
    q = torch.div((idx - cut).clamp_min(0), max(base, 1), rounding_mode="floor")
    owner = torch.where(small, own_a, rem + q)
    local = torch.where(small, idx - own_a * (base + 1), idx - cut - q * base)
    return owner, local


def local_row_index(all_idx, lo, hi):
    """Map global row indices to this owner's local rows; rows owned by other ranks map to the sentinel
    (hi - lo).  Returns (local_idx int64, owned bool)."""
    owned = (all_idx >= lo) & (all_idx < hi)
    local = torch.where(owned, all_idx - lo, torch.full_like(all_idx, hi - lo))
    return local, owned


def route_to_owners(idx, n_rows, world):
    """Stable bucketing of a rank's requests by owner: -> (order, counts) with idx[order] grouped by owner rank (request
    order kept inside a group) and counts[w] = requests for owner w (the all-to-all split sizes)."""
    owner, _ = owner_of_rows(idx, n_rows, world)
    order = torch.sort(owner, stable=True).indices
    counts = torch.bincount(owner, minlength=world)
    return order, counts


class _RoutedRowsFn(torch.autograd.Function):
    """rows[b] = table[u[b]] for a block-sharded table with torch.distributed collectives only (any backend): SURVEY 8e row 2
    as written -- all-to-all of the requested indices to their owners, owner-side gather of exactly the requested rows,
    all-to-all of the rows back; backward routes (index, gradient row) pairs to the owners the same way and the owner
    segment-sums them into its dense shard gradient (already the GLOBAL sum: not all-reduced again).  Variable split sizes
    cost one host read of `world` counts per step; the NVLink peer-memory transport (_PeerUserTowerFn) needs neither."""

    @staticmethod
    def forward(ctx, u, shard, n_rows, lo, hi, group, gather_fn, scatter_fn):
        world = dist.get_world_size(group)
        u = u.contiguous().view(-1)
        order, counts = route_to_owners(u, n_rows, world)
        send_counts = counts.tolist()
        recv_counts_t = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts_t, counts, group=group)
        recv_counts = recv_counts_t.tolist()
        req = torch.empty(sum(recv_counts), dtype=torch.int64, device=u.device)
        dist.all_to_all_single(req, u[order].contiguous(), recv_counts, send_counts, group=group)     # indices -> owners
        rows_out = gather_fn(shard, req - lo)                                                        # only owned rows
        back = torch.empty(u.numel(), shard.shape[1], dtype=shard.dtype, device=u.device)
        dist.all_to_all_single(back, rows_out.contiguous(), send_counts, recv_counts, group=group)   # rows -> requesters
        rows = torch.empty_like(back)
        rows[order] = back
        ctx.save_for_backward(order, req)
        ctx.meta = (lo, hi, group, send_counts, recv_counts, scatter_fn)
        return rows

    @staticmethod
    def backward(ctx, grows):
        order, req = ctx.saved_tensors
        lo, hi, group, send_counts, recv_counts, scatter_fn = ctx.meta
        g_in = torch.empty(req.numel(), grows.shape[1], dtype=grows.dtype, device=grows.device)
        dist.all_to_all_single(g_in, grows.contiguous()[order].contiguous(), recv_counts, send_counts, group=group)
        # NOTE the order of summation: requests arrive grouped by requesting rank, in request order -> deterministic
        gshard = scatter_fn(req - lo, g_in, hi - lo)
        return None, gshard, None, None, None, None, None, None


class _PeerUserTowerFn(torch.autograd.Function):
    """u_f = linear2(relu(linear1(relu(table[u])))) for a table block-sharded in NVLink peer (symmetric) memory.
    forward : barrier (every owner finished its previous optimizer step), then ONE gather kernel whose row loads go over
              NVLink to the owning GPU -- no index exchange, no row all-to-all;
    backward: the MLP data gradient (masked by the gather's ReLU) is written into this rank's exchange slot, one kernel
              publishes the indices + barriers + collects all index lists, the owner sums the gradient rows it owns from
              the peers' slots in (rank, position) order into the dense shard gradient."""

    @staticmethod
    def forward(ctx, u, shard, w1, b1, w2, b2, table):
        from . import _lib as L
        u = u.contiguous().view(-1)
        B, E, F = u.numel(), shard.shape[1], w2.shape[0]
        dev, st = shard.device, L.stream()
        f32 = dict(dtype=torch.float32, device=dev)
        h0, h1, out = torch.empty(B, E, **f32), torch.empty(B, E, **f32), torch.empty(B, F, **f32)
        table.xch.barrier(0)
        L.call("dcue_peer_gather_relu_fwd", table.shard_ptrs_dev, table.user_count, table.world_size, u.data_ptr(), B, E,
               h0.data_ptr(), None, table.err_flag().data_ptr(), st)
        L.call("dcue_linear_fwd", h0.data_ptr(), E, w1.data_ptr(), b1.data_ptr(), B, E, E, 1, h1.data_ptr(), E, st)
        L.call("dcue_linear_fwd", h1.data_ptr(), E, w2.data_ptr(), b2.data_ptr(), B, E, F, 0, out.data_ptr(), F, st)
        ctx.save_for_backward(u, w1, w2, h0, h1)
        ctx.table = table
        return out

    @staticmethod
    def backward(ctx, gout):
        from . import _lib as L
        u, w1, w2, h0, h1 = ctx.saved_tensors
        table = ctx.table
        B, E, F = u.numel(), h0.shape[1], w2.shape[0]
        dev, st = h0.device, L.stream()
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
        xch = table.xch
        # single exchange slot: the forward barrier of this step came after every peer's previous scatter
        L.call("dcue_linear_dgrad", dh1.data_ptr(), E, w1.data_ptr(), B, E, E, h0.data_ptr(), E, xch.rows.data_ptr(), E, st)
        gshard = xch.scatter_add(xch.exchange_indices(u), B, table.lo, table.hi)
        return None, gshard, gw1, gb1, gw2, gb2, None


class ShardedUserTable(torch.nn.Module):
    """Row-sharded replacement for UserEmbeddings' lookup table (BASELINE cfg4: 1M users x 300 over 8 GPUs).
    Each rank owns rows [lo, hi) (+ its optimizer state); the MLP stays replicated / data parallel.

    transport="peer" (default on CUDA + NCCL): the shard lives in symmetric memory and rows / gradient rows move by
    NVLink loads inside the gather / scatter kernels (csrc/peer.cu).  transport="collective": torch.distributed
    all-to-all of indices and rows (any backend; also what the CPU/gloo tests drive)."""

    def __init__(self, user_embd, group=None, transport=None, capacity=None, gather_fn=None, scatter_fn=None):
        super().__init__()
        from . import ops
        self._ops = ops
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        full = user_embd.embeddings.weight.detach()
        self.user_count, E = full.shape
        self.lo, self.hi = shard_rows(self.user_count, self.rank, self.world_size)
        if transport is None:
            transport = "peer" if (full.is_cuda and dist.get_backend(group) == "nccl"
                                   and os.environ.get("DCUE_SHARD_TRANSPORT", "peer") == "peer") else "collective"
        self.transport = transport
        self.linear1, self.linear2 = user_embd.linear1, user_embd.linear2
        self._gather_fn = gather_fn or ops.gather_rows
        self._scatter_fn = scatter_fn or (lambda idx, rows, n: ops.scatter_rows(idx, rows, n))
        self._err = None
        self.xch = None
        if transport == "peer":
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            # every rank allocates the same size (symmetric): the largest shard
            rows_max = self.user_count // self.world_size + (1 if self.user_count % self.world_size else 0)
            buf = symm.empty(rows_max * E, dtype=torch.float32, device=full.device)
            self._shard_hdl = symm.rendezvous(buf, pg.group_name)
            self.shard_ptrs_dev = self._shard_hdl.buffer_ptrs_dev
            shard = buf[: (self.hi - self.lo) * E].view(self.hi - self.lo, E)
            shard.copy_(full[self.lo:self.hi])
            self.shard = torch.nn.Parameter(shard)
            self._capacity = capacity
        else:
            self.shard = torch.nn.Parameter(full[self.lo:self.hi].clone())

    def err_flag(self):
        if self._err is None or self._err.device != self.shard.device:
            self._err = torch.zeros(1, dtype=torch.int32, device=self.shard.device)
        return self._err

    _err_flag = err_flag

    def forward(self, user_idx):
        shape = user_idx.shape
        u = user_idx.reshape(-1).to(self.shard.device)
        if self.transport == "peer":
            if self.xch is None or self.xch.capacity < u.numel():
                self.xch = PeerExchange(self.group, self.shard.device, max(u.numel(), self._capacity or 0), self.shard.shape[1])
            out = _PeerUserTowerFn.apply(u, self.shard, self.linear1.weight, self.linear1.bias, self.linear2.weight,
                                         self.linear2.bias, self)
        else:
            rows = _RoutedRowsFn.apply(u, self.shard, self.user_count, self.lo, self.hi, self.group, self._gather_fn,
                                       self._scatter_fn)
            out = self._ops.UserMLPFn.apply(rows, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
        return out.view(*shape, -1)

    def raise_if_index_error(self):
        if self._err is not None and int(self._err.item()):
            self._err.zero_()
            raise IndexError("index out of range in self")
        if self.xch is not None:
            self.xch.check()

    def gather_full_table(self):
        """[U,E] table on every rank (for checkpoints in the reference's state_dict layout)."""
        sizes = [shard_rows(self.user_count, r, self.world_size) for r in range(self.world_size)]
        m = max(h - l for l, h in sizes)
        pad = torch.zeros(m, self.shard.shape[1], dtype=self.shard.dtype, device=self.shard.device)
        pad[: self.hi - self.lo] = self.shard.detach()
        out = [torch.empty_like(pad) for _ in range(self.world_size)]
        dist.all_gather(out, pad, group=self.group)
        return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)])


def shard_user_table(model, group=None, transport=None, capacity=None):
    """Replace model.user_embd by its row-sharded version (call before building the optimizer)."""
    model.user_embd = ShardedUserTable(model.user_embd, group, transport, capacity)
    return model


# ------------------------------------------------------------------------------ song-sharded eval (cfg5)
def user_block(n_users, rank, world):
    """Users whose merged top-k lives on `rank`: contiguous block [lo, hi) of shard_slice."""
    return shard_slice(n_users, rank, world)


def exchange_topk_by_user_block(s, i, group=None):
    """Per-rank top-k lists [U,k] (this rank's songs) -> [world, U_r, k] lists of the users in THIS rank's block, one part
    per song shard (all-to-all by user block, SURVEY 8e row 3: each rank receives W x U/W x k pairs instead of the
    W x U x k of an all-gather)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    U, k = s.shape
    sizes = [b - a for a, b in (user_block(U, r, world) for r in range(world))]
    mine = sizes[rank]
    s_parts = torch.empty(world, mine, k, dtype=s.dtype, device=s.device)
    i_parts = torch.empty(world, mine, k, dtype=i.dtype, device=i.device)
    send = [n * k for n in sizes]
    recv = [mine * k] * world
    dist.all_to_all_single(s_parts.view(-1), s.contiguous().view(-1), recv, send, group=group)
    dist.all_to_all_single(i_parts.view(-1), i.contiguous().view(-1), recv, send, group=group)
    return s_parts, i_parts


def sharded_topk(user_factors, item_factors_local, k, item_offset, group=None, gather=False, user_tile=None):
    """Song-sharded eval (BASELINE cfg5): every rank scores all users against ITS songs (fused score GEMM + top-k), the
    per-rank lists are exchanged all-to-all BY USER BLOCK and every rank merges only its own U/W users.
    -> (scores [U_r,k], song idx [U_r,k], (lo, hi)) for this rank's user block; gather=True all-gathers the merged blocks so
    that every rank returns the full [U,k] (only when the caller needs it: the merged result is 8x smaller than the
    exchange).  With user_tile the users are processed in tiles so that the exchange of tile t overlaps the scoring of
    tile t+1 (the all-to-all runs on NCCL's stream)."""
    from . import eval as ev
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        s, i = ev.topk_scores(user_factors, item_factors_local, k, item_offset=item_offset)
        return (s, i) if gather else (s, i, (0, user_factors.shape[0]))
    rank = dist.get_rank(group)
    U = user_factors.shape[0]
    lo, hi = user_block(U, rank, world)
    packed_items = ev.normalize_factors(item_factors_local)
    if user_tile is None:
        user_tile = int(os.environ.get("DCUE_EVAL_USER_TILE", "0")) or U
    # a tile = the same sub-range of every rank's block, so that each tile's exchange is itself an all-to-all by block
    n_tiles = max(1, -(-U // user_tile))
    out_s = torch.empty(hi - lo, k, dtype=torch.float32, device=user_factors.device)
    out_i = torch.empty(hi - lo, k, dtype=torch.int64, device=user_factors.device)
    blocks = [user_block(U, r, world) for r in range(world)]
    pending = None

    def finish(p):
        works, s_parts, i_parts, (x, y), _keep = p
        for w in works:
            w.wait()                      # stream-side wait: the merge kernel follows the exchange
        ms, mi = ev.merge_topk_parts(s_parts, i_parts)
        out_s[x:y], out_i[x:y] = ms, mi

    n_items_local = item_factors_local.shape[0]
    use_global = (os.environ.get("DCUE_EVAL_GLOBAL_SEED", "0") != "0"
                  and int(ev.L.lib().dcue_topk_sample_r(max(1, U // n_tiles), n_items_local, k)) > 0)
    failed_rows = []

    for t in range(n_tiles):
        sub = [shard_slice(b - a, t, n_tiles) for a, b in blocks]           # sub-range of every block
        if n_tiles == 1:
            uf = user_factors
        else:
            rows = torch.cat([torch.arange(a + x, a + y, device=user_factors.device) for (a, _), (x, y) in zip(blocks, sub)])
            uf = user_factors[rows]
        sizes = [y - x for x, y in sub]
        mine = sizes[rank]
        nu_t = uf.shape[0]
        samp = None
        if use_global:
            # global-threshold protocol: the r best SAMPLE scores of every shard are merged per user (all-to-all by user block),
            # the r-th best of the union is a threshold that ~4k songs of all shards TOGETHER exceed, so every shard keeps ~4k/W
            # candidates per user instead of 4k (fewer appends, no selection, shorter sort) and the final merge sees ~4k entries
            upk = ev.normalize_factors(uf)
            samp = ev.sample_scores(upk, nu_t, packed_items, n_items_local, k)
        if samp is not None:
            r = samp.shape[1]
            sp = torch.empty(world, mine, r, dtype=torch.float32, device=samp.device)
            dist.all_to_all_single(sp.view(-1), samp.view(-1), [mine * r] * world, [n * r for n in sizes], group=group)
            dummy = torch.zeros(world, mine, r, dtype=torch.int64, device=samp.device)
            merged, _ = ev.merge_topk_parts(sp, dummy)
            m = max(sizes)
            thr_pad = torch.full((m,), float("-inf"), dtype=torch.float32, device=samp.device)
            thr_pad[:mine] = merged[:, r - 1]
            thr_all = torch.empty(world, m, dtype=torch.float32, device=samp.device)
            dist.all_gather_into_tensor(thr_all.view(-1), thr_pad, group=group)
            thr = torch.cat([thr_all[q, :n] for q, n in enumerate(sizes)])
            s, i = ev.topk_scores_seeded(upk, nu_t, packed_items, n_items_local, k, thr, item_offset=item_offset)
        else:
            s, i = ev.topk_scores(uf, item_factors_local, k, item_offset=item_offset, normalized_items=packed_items)
        s_parts = torch.empty(world, mine, k, dtype=s.dtype, device=s.device)
        i_parts = torch.empty(world, mine, k, dtype=i.dtype, device=i.device)
        send, recv = [n * k for n in sizes], [mine * k] * world
        # asynchronous: the exchange of tile t runs on NCCL's stream while the scorer works on tile t+1
        works = [dist.all_to_all_single(s_parts.view(-1), s.view(-1), recv, send, group=group, async_op=True),
                 dist.all_to_all_single(i_parts.view(-1), i.view(-1), recv, send, group=group, async_op=True)]
        if pending is not None:
            finish(pending)
        pending = (works, s_parts, i_parts, sub[rank], (s, i))
    finish(pending)
    if use_global:
        # a user whose merged list is shorter than k: the guessed threshold was too high (probability ~1e-4 per user).
        # The owners publish those users, every shard scores them exactly (thresholds from -inf) and the owner merges again.
        bad = torch.nonzero(out_i[:, k - 1] < 0).flatten() + lo                   # global user ids (host sync: eval only)
        cnt = torch.tensor([bad.numel()], dtype=torch.int64, device=bad.device)
        cnts = torch.empty(world, dtype=torch.int64, device=bad.device)
        dist.all_gather_into_tensor(cnts, cnt, group=group)
        cl = cnts.tolist()
        if sum(cl):
            mx = max(cl)
            pad = torch.full((mx,), -1, dtype=torch.int64, device=bad.device)
            pad[: bad.numel()] = bad
            allbad = torch.empty(world, mx, dtype=torch.int64, device=bad.device)
            dist.all_gather_into_tensor(allbad.view(-1), pad, group=group)
            ids = torch.cat([allbad[q, :n] for q, n in enumerate(cl)])
            es, ei = ev._topk_exact(user_factors[ids].contiguous(), packed_items[0], packed_items[1], n_items_local, k, item_offset)
            gs = torch.empty(world, ids.numel(), k, dtype=es.dtype, device=es.device)
            gi = torch.empty(world, ids.numel(), k, dtype=ei.dtype, device=ei.device)
            dist.all_gather_into_tensor(gs.view(-1), es.contiguous().view(-1), group=group)
            dist.all_gather_into_tensor(gi.view(-1), ei.contiguous().view(-1), group=group)
            off = sum(cl[:rank])
            if cl[rank]:
                ms, mi = ev.merge_topk_parts(gs[:, off: off + cl[rank]].contiguous(), gi[:, off: off + cl[rank]].contiguous())
                out_s[bad - lo], out_i[bad - lo] = ms, mi
    if not gather:
        return out_s, out_i, (lo, hi)
    m = max(b - a for a, b in blocks)
    pad_s = torch.zeros(m, k, dtype=out_s.dtype, device=out_s.device)
    pad_i = torch.zeros(m, k, dtype=out_i.dtype, device=out_i.device)
    pad_s[: hi - lo], pad_i[: hi - lo] = out_s, out_i
    all_s = torch.empty(world, m, k, dtype=out_s.dtype, device=out_s.device)
    all_i = torch.empty(world, m, k, dtype=out_i.dtype, device=out_i.device)
    dist.all_gather_into_tensor(all_s.view(-1), pad_s.view(-1), group=group)
    dist.all_gather_into_tensor(all_i.view(-1), pad_i.view(-1), group=group)
    return (torch.cat([all_s[r, : b - a] for r, (a, b) in enumerate(blocks)]),
            torch.cat([all_i[r, : b - a] for r, (a, b) in enumerate(blocks)]))


_EVAL_GROUPS = {}


def eval_layout(song_shards, group=None):
    """2-D layout of the eval over the ranks: `song_shards` song shards x (world / song_shards) user groups.  Rank r scores the
    users of group r // song_shards against song shard r % song_shards; the ranks of one user group merge their lists with
    sharded_topk(group=subgroup).  song_shards == world is BASELINE cfg5's pure song sharding; fewer song shards mean longer
    song streams per work item (the scorer's per-user-tile start / finish cost is amortised over more songs) at the price of
    every rank holding I / song_shards song factors.  Collective: every rank must call it with the same arguments.
    -> dict(song_shard, user_group, n_user_groups, group)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world % song_shards:
        raise ValueError("song_shards must divide the world size")
    key = (song_shards, world, id(group))
    if key not in _EVAL_GROUPS:
        subs = []
        for ug in range(world // song_shards):
            ranks = [ug * song_shards + j for j in range(song_shards)]
            subs.append(dist.new_group(ranks) if song_shards < world else (group if group is not None else dist.group.WORLD))
        _EVAL_GROUPS[key] = subs
    ug = rank // song_shards
    return dict(song_shard=rank % song_shards, user_group=ug, n_user_groups=world // song_shards, group=_EVAL_GROUPS[key][ug])
