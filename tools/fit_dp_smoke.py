"""DCUE.fit under torchrun (the trainer itself drives the data-parallel path): every rank trains on its shard of the
synthetic world, rank 0 writes the checkpoints, close() tears the process group down normally.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/fit_dp_smoke.py
Prints one JSON line on rank 0: epochs run, parameters identical across ranks, checkpoint files."""
import importlib
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synthetic_data import SynthItemSet, SynthPredSet, SynthTrainSet, SynthWorld  # noqa: E402

pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = SynthWorld(n_users=32, n_songs=40)
    tr, va = SynthTrainSet(w, w.pairs[:160]), SynthTrainSet(w, w.pairs[160:])
    t = pkg.DCUE(batch_size=4, neg_batch_size=w.negs, lr=1e-4, num_epochs=1, eval_pct=1.0)
    t.num_workers = 0
    np.random.seed(0)
    torch.manual_seed(100 + rank)          # different initialisation per rank: DataParallelDCUE must broadcast rank 0's
    d = tempfile.mkdtemp() if rank == 0 else "/nonexistent-on-purpose"
    t.fit(tr, va, va, SynthPredSet(w, w.pairs[160:]), SynthPredSet(w, w.pairs[:160]), SynthItemSet(w), w.n_users, w.n_songs,
          "triplets.txt", "metadata.csv", d)
    # every rank must hold the same parameters after training
    flat = torch.cat([p.detach().flatten() for p in t.model.parameters()])
    ref = flat.clone()
    if world > 1:
        dist.broadcast(ref, 0)
    diff = torch.tensor([(flat - ref).abs().max().item()], device=flat.device)
    if world > 1:
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    files = []
    if rank == 0:
        for root, _, fs in os.walk(d):
            files += fs
        print(json.dumps({"world": world, "nn_epoch": t.nn_epoch, "dp_wrapped": t._dp is not None,
                          "max_param_diff_between_ranks": diff.item(), "checkpoints": sorted(files)[:4],
                          "finite": bool(torch.isfinite(flat).all())}), flush=True)
    t.close()


if __name__ == "__main__":
    main()
