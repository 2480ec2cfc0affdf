"""Song-sharded eval (BASELINE cfg5) on real GPUs, one rank per GPU (torchrun): every rank scores all users
against ITS shard of the song factors, the per-rank top-k lists are all-gathered and merged; the result must
equal the single-GPU top-k over all songs.  Also prints the sharded throughput.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/sharded_eval_parity.py [n_users] [n_items] [k]"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    nu = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    ni = int(sys.argv[2]) if len(sys.argv) > 2 else 50000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    check = nu <= 65536
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    par = importlib.import_module("amplifai-deepcontentrecommenders_b200.parallel")
    g = torch.Generator(device=dev).manual_seed(3)          # same factors on every rank
    uf = torch.randn(nu, 100, generator=g, device=dev)
    itf = torch.randn(ni, 100, generator=g, device=dev)
    lo, hi = par.shard_slice(ni, rank, world)
    for _ in range(2):
        par.sharded_topk(uf[:1024], itf[lo:hi], k, lo)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s, i = par.sharded_topk(uf, itf[lo:hi], k, lo)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ok = True
    msg = ""
    if check:
        s1, i1 = pkg.eval.topk_scores(uf, itf, k)
        same_scores = (s - s1).abs().max().item()
        agree = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i.cpu(), i1.cpu())) / (nu * k)
        ok = same_scores < 1e-6 and agree > 0.9999
        msg = "max|score diff| %.1e set agreement %.6f" % (same_scores, agree)
    if rank == 0:
        print("SHARDED_EVAL world=%d users=%d songs=%d k=%d: %.1f ms -> %.3g users/s %s %s"
              % (world, nu, ni, k, ms.item(), nu / (ms.item() * 1e-3), msg, "PASS" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
