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
