"""Data-parallel parity on real GPUs (launch with torchrun, one rank per GPU):
the global batch split across ranks (BatchNorm statistics all-reduced, flat gradient all-reduce)
must reproduce the single-process step on the full batch.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py"""
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
    ok = True
    for mt in ("truedcuemel1dbn", "truedcuemel1dres"):
        B, N, U = 16 * world, 4, 60   # reference S = B*(1+N) >= 160: more 128-row tiles than SMs at layer 1
        params = fixtures.make_params(mt, seed=0, user_count=U)
        u, pos, neg = fixtures.make_inputs(B, N, U, seed=1)
        cfg = {"feature_dim": 100, "conv_hidden": 128, "user_embdim": 300, "user_count": U, "model_type": mt}
        # ---- single process, full batch (every rank computes it; cheap)
        ref = pkg.DCUENet(cfg)
        ref.load_state_dict(params)
        ref = ref.to(dev).train()
        loss_ref = ref.hinge_loss_step(u.to(dev), pos.to(dev), neg.to(dev), 0.2)
        loss_ref.backward()
        # ---- data parallel
        net = pkg.DCUENet(cfg)
        net.load_state_dict(params)
        net = net.to(dev).train()
        dp = par.DataParallelDCUE(net)
        lo, hi = par.shard_slice(B, rank, world)
        loss = dp.loss_step(u[lo:hi].to(dev), pos[lo:hi].to(dev), neg[lo:hi].to(dev), 0.2)
        loss.backward()
        dp.reduce_gradients()
        total = dp.reduce_loss(loss)
        worst, errs = 0.0, []
        for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            e = ((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-30)).item()
            errs.append((e, k, q.grad.norm().item()))
            worst = max(worst, e)
        if rank == 0 and worst > 1e-3:
            print("DP_PARITY_DETAIL", mt, sorted(errs, reverse=True)[:6], flush=True)
        berr = max(((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()
                   for (_, a), (_, b) in zip(net.named_buffers(), ref.named_buffers())) if list(net.named_buffers()) else 0.0
        lerr = abs(total.item() - loss_ref.item()) / abs(loss_ref.item())
        if rank == 0:
            print("DP_PARITY %s world=%d loss_rel=%.2e worst_grad_l2=%.2e buffers=%.2e" % (mt, world, lerr, worst, berr), flush=True)
        ok = ok and lerr < 1e-5 and worst < 2e-2 and berr < 1e-5
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DP_PARITY", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
