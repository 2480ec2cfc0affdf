"""Row-sharded user table (BASELINE cfg4) on real GPUs, one rank per GPU (torchrun): the sharded lookup +
exchange must reproduce the replicated table: same loss, same gradients (shard gradient == the owner's rows of
the dense table gradient), and after an Adam step the gathered table equals the replicated one.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/sharded_table_parity.py"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    par = importlib.import_module("amplifai-deepcontentrecommenders_b200.parallel")
    mt, U, N = "truedcuemel1dbn", 1003, 4
    B = 8 * world
    params = fixtures.make_params(mt, seed=0, user_count=U)
    u, pos, neg = fixtures.make_inputs(B, N, U, seed=1, zipf=True)
    cfg = {"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": U, "model_type": mt}
    lo_b, hi_b = par.shard_slice(B, rank, world)

    def run(sharded):
        net = pkg.DCUENet(cfg)
        net.load_state_dict(params)
        net = net.to(dev).train()
        if sharded:
            par.shard_user_table(net)
        dp = par.DataParallelDCUE(net, broadcast=False)
        opt = torch.optim.Adam(net.parameters(), 1e-3)
        loss = dp.loss_step(u[lo_b:hi_b].to(dev), pos[lo_b:hi_b].to(dev), neg[lo_b:hi_b].to(dev), 0.2)
        loss.backward()
        dp.reduce_gradients()
        tot = dp.reduce_loss(loss).item()
        if sharded:
            g_table = net.user_embd.shard.grad.clone()
        else:
            g_table = net.user_embd.embeddings.weight.grad.clone()
        g_l1 = net.user_embd.linear1.weight.grad.clone()
        opt.step()
        table = net.user_embd.gather_full_table() if sharded else net.user_embd.embeddings.weight.detach().clone()
        return tot, g_table, g_l1, table

    l0, gt0, gl0, t0 = run(False)
    l1, gt1, gl1, t1 = run(True)
    lo, hi = par.shard_rows(U, rank, world)
    e_loss = abs(l0 - l1) / abs(l0)
    e_gt = ((gt1 - gt0[lo:hi]).abs().max() / gt0.abs().max()).item()
    e_gl = ((gl1 - gl0).abs().max() / gl0.abs().max()).item()
    e_t = (t1 - t0).abs().max().item()
    ok = e_loss < 1e-6 and e_gt < 1e-5 and e_gl < 1e-5 and e_t < 1e-6
    print("SHARDED_TABLE rank=%d world=%d loss_rel=%.2e shard_grad=%.2e mlp_grad=%.2e table_after_adam=%.2e %s"
          % (rank, world, e_loss, e_gt, e_gl, e_t, "PASS" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
