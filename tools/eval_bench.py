"""Eval scorer throughput (BASELINE cfg5 metric: scored users/sec, top-k): all-pairs cosine scores
of user factors against song factors with the fused top-100.
  python tools/eval_bench.py [n_users] [n_items] [k]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    nu = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    ni = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    ev, L = pkg.eval, pkg._lib
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    uf = torch.randn(nu, 100, generator=g, device=dev)
    itf = torch.randn(ni, 100, generator=g, device=dev)
    items = ev.normalize_factors(itf)
    for _ in range(2):
        ev.topk_scores(uf[:4096], itf, k, normalized_items=items)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s, i = ev.topk_scores(uf, itf, k, normalized_items=items)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flops = 2.0 * 100 * nu * ni
    # spot check against fp32 torch on a few users
    chk = torch.topk(torch.nn.functional.normalize(uf[:64]) @ torch.nn.functional.normalize(itf).T, k, dim=1)
    err = (s[:64] - chk.values).abs().max().item()
    agree = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(i[:64].cpu(), chk.indices.cpu())) / (64 * k)
    print(json.dumps({"metric": "eval scored users/sec (top-%d)" % k, "value": nu / (ms * 1e-3), "unit": "users/s", "ms": ms,
                      "n_users": nu, "n_items": ni, "tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
                      "max_score_err_vs_fp32": err, "topk_set_agreement_vs_fp32": agree}))


if __name__ == "__main__":
    main()
