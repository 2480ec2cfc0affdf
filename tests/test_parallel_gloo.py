"""Host-side multi-rank logic on CPU (gloo, world_size 2): batch sharding, the gradient bucket
selection and the all-reduce plumbing of parallel.DataParallelDCUE (no kernels run here)."""
import importlib
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

par = importlib.import_module("amplifai-deepcontentrecommenders_b200.parallel")


def test_shard_slice_covers_everything():
    for n in (8192, 10, 7):
        for w in (1, 2, 3, 8):
            spans = [par.shard_slice(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_bucket_excludes_batchnorm_affine():
    names = [("conv.bn0.weight", 0), ("conv.layer1.weight", 0), ("conv.bn3.bias", 0), ("user_embd.embeddings.weight", 0),
             ("conv.fc.bias", 0)]
    # bn1..bn5 affine gradients come from all-reduced sums; the replicated table's dense gradient is built from the
    # all-gathered gradient rows of every rank (ops.UserTowerFn): neither goes through the flat SUM all-reduce
    assert par.flat_bucket_names(names) == ["conv.bn0.weight", "conv.layer1.weight", "conv.fc.bias"]


def test_sharded_table_index_bookkeeping():
    U, W = 1000, 8
    spans = [par.shard_rows(U, r, W) for r in range(W)]
    idx = torch.tensor([0, 124, 125, 999, 500, 125])
    owners = torch.zeros_like(idx)
    for r, (lo, hi) in enumerate(spans):
        local, owned = par.local_row_index(idx, lo, hi)
        assert (local[owned] == idx[owned] - lo).all() and (local[~owned] == hi - lo).all()   # sentinel = shard size
        owners += owned.long()
    assert (owners == 1).all()                                     # every index has exactly one owner
    assert par.flat_bucket_names([("user_embd.shard", 0), ("user_embd.linear1.weight", 0)]) == ["user_embd.linear1.weight"]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    torch.manual_seed(rank)  # different init per rank: broadcast must equalise it
    net = pkg.DCUENet({"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": 20,
                       "model_type": "truedcuemel1dbn"})
    dp = par.DataParallelDCUE(net)
    w = net.conv.layer1.weight.detach().clone()
    # fake local gradients: rank-dependent constants
    for n, p in net.named_parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    dp.reduce_gradients()
    stats = torch.tensor([1.0 + rank, 2.0], dtype=torch.float64)
    dp.all_reduce_sum(stats)
    loss = dp.reduce_loss(torch.tensor(0.25 * (rank + 1)))
    rows = dp.all_gather_rows(torch.full((2, 3), float(rank)))
    res = dict(w=w, g_conv=net.conv.layer1.weight.grad[0, 0, 0].item(), g_bn=net.conv.bn1.weight.grad[0].item(),
               g_tab=net.user_embd.embeddings.weight.grad[3, 7].item(), stats=stats, loss=loss.item(),
               hooked=net.conv._dp is dp and net.user_embd._dp is dp, rows=rows)
    torch.save(res, out % rank)
    dist.destroy_process_group()


def test_data_parallel_plumbing_world2(tmp_path):
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out % 0), torch.load(out % 1)
    assert torch.equal(r0["w"], r1["w"])                      # parameters broadcast from rank 0
    assert r0["g_conv"] == r1["g_conv"] == 3.0                # 1 + 2 summed
    assert r0["g_tab"] == 1.0 and r1["g_tab"] == 2.0          # table gradient: exchanged as rows in the backward, not here
    assert torch.equal(r0["rows"], torch.tensor([[0.0] * 3] * 2 + [[1.0] * 3] * 2)) and torch.equal(r0["rows"], r1["rows"])
    assert r0["g_bn"] == 1.0 and r1["g_bn"] == 2.0            # BN affine grads are already global: untouched
    assert torch.equal(r0["stats"], torch.tensor([3.0, 4.0], dtype=torch.float64))
    assert abs(r0["loss"] - 0.75) < 1e-12 and r0["hooked"] and r1["hooked"]


# ------------------------------------------------------------------ round 2: routed sharded table + eval exchange
def test_owner_arithmetic_and_routing():
    for U, W in ((1000, 8), (10, 3), (7, 8), (1000003, 8)):
        idx = torch.arange(U) if U < 10000 else torch.randint(0, U, (4000,), generator=torch.Generator().manual_seed(1))
        owner, local = par.owner_of_rows(idx, U, W)
        for r in range(W):
            lo, hi = par.shard_rows(U, r, W)
            m = (idx >= lo) & (idx < hi)
            assert (owner[m] == r).all() and (local[m] == idx[m] - lo).all()
    idx = torch.tensor([999, 0, 500, 1, 999, 124, 125])
    order, counts = par.route_to_owners(idx, 1000, 8)
    owner, _ = par.owner_of_rows(idx, 1000, 8)
    assert counts.tolist() == torch.bincount(owner, minlength=8).tolist() and counts.sum() == idx.numel()
    assert (owner[order][1:] >= owner[order][:-1]).all()
    # stable: requests for one owner keep their request order (determinism of the owner-side sums)
    for w in range(8):
        pos = order[owner[order] == w]
        assert (pos[1:] > pos[:-1]).all()


def _table_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    U, E, B = 37, 6, 11
    g = torch.Generator().manual_seed(0)
    full = torch.randn(U, E, generator=g)

    class _Emb(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embeddings = torch.nn.Embedding(U, E)
            self.embeddings.weight.data.copy_(full)
            self.linear1, self.linear2 = torch.nn.Linear(E, E), torch.nn.Linear(E, 3)

    def scatter(idx, rows, n):
        o = torch.zeros(n, rows.shape[1])
        o.index_add_(0, idx, rows)
        return o

    tab = par.ShardedUserTable(_Emb(), transport="collective", gather_fn=lambda sh, i: sh[i], scatter_fn=scatter)
    u = torch.randint(0, U, (B,), generator=torch.Generator().manual_seed(10 + rank))
    u[0] = 5                      # a row requested by both ranks
    rows = par._RoutedRowsFn.apply(u, tab.shard, U, tab.lo, tab.hi, None, tab._gather_fn, tab._scatter_fn)
    gr = torch.randn(B, E, generator=torch.Generator().manual_seed(20 + rank))
    rows.backward(gr)
    torch.save(dict(u=u, rows=rows.detach(), gr=gr, gshard=tab.shard.grad, lo=tab.lo, hi=tab.hi, full=full,
                    table=tab.gather_full_table()), out % rank)
    dist.destroy_process_group()


def test_routed_sharded_table_world2(tmp_path):
    port = 31500 + os.getpid() % 2000
    out = str(tmp_path / "t%d.pt")
    mp.spawn(_table_worker, args=(2, port, out), nprocs=2, join=True)
    r = [torch.load(out % k) for k in range(2)]
    full = r[0]["full"]
    ref = torch.zeros_like(full)
    for k in range(2):
        assert torch.equal(r[k]["rows"], full[r[k]["u"]])                     # forward == plain lookup
        assert torch.equal(r[k]["table"], full)
        ref.index_add_(0, r[k]["u"], r[k]["gr"])
    for k in range(2):
        assert torch.allclose(r[k]["gshard"], ref[r[k]["lo"]:r[k]["hi"]], atol=1e-6)   # owner holds the GLOBAL row sums


def _topk_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    U, k = 7, 3
    s = torch.arange(U * k, dtype=torch.float32).view(U, k) + 100 * rank
    i = torch.arange(U * k, dtype=torch.int64).view(U, k) + 1000 * rank
    sp, ip = par.exchange_topk_by_user_block(s, i)
    torch.save(dict(sp=sp, ip=ip, blk=par.user_block(U, rank, world)), out % rank)
    dist.destroy_process_group()


def test_topk_exchange_by_user_block_world2(tmp_path):
    port = 33500 + os.getpid() % 2000
    out = str(tmp_path / "k%d.pt")
    mp.spawn(_topk_worker, args=(2, port, out), nprocs=2, join=True)
    U, k = 7, 3
    for rank in range(2):
        r = torch.load(out % rank)
        lo, hi = r["blk"]
        assert r["sp"].shape == (2, hi - lo, k)
        for src in range(2):       # part `src` = what rank `src` computed for MY users
            assert torch.equal(r["sp"][src], torch.arange(U * k, dtype=torch.float32).view(U, k)[lo:hi] + 100 * src)
            assert torch.equal(r["ip"][src], torch.arange(U * k, dtype=torch.int64).view(U, k)[lo:hi] + 1000 * src)


def _layout_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lay = par.eval_layout(2)
    t = torch.tensor([float(rank)])
    dist.all_reduce(t, group=lay["group"])            # sums the ranks of my user group only
    full = par.eval_layout(world)                     # pure song sharding: the whole world is one group
    t2 = torch.tensor([1.0])
    dist.all_reduce(t2, group=full["group"])
    torch.save(dict(lay={k: v for k, v in lay.items() if k != "group"}, s=t.item(), s2=t2.item(),
                    full={k: v for k, v in full.items() if k != "group"}), out % rank)
    dist.destroy_process_group()


def test_eval_layout_world4(tmp_path):
    port = 35500 + os.getpid() % 2000
    out = str(tmp_path / "l%d.pt")
    mp.spawn(_layout_worker, args=(4, port, out), nprocs=4, join=True)
    for rank in range(4):
        r = torch.load(out % rank)
        assert r["lay"] == dict(song_shard=rank % 2, user_group=rank // 2, n_user_groups=2)
        assert r["s"] == (1.0 if rank < 2 else 5.0)          # ranks {0,1} and {2,3}
        assert r["full"] == dict(song_shard=rank, user_group=0, n_user_groups=1) and r["s2"] == 4.0
