"""Diagnostic: host and device time of every graph-replayed training step around a device synchronisation (and around the
fork of a helper process), to find one-off stalls that a mean over K steps hides.  usage: python tools/step_probe.py [fork]"""
import importlib
import os
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
    optim = importlib.import_module("amplifai-deepcontentrecommenders_b200.optim")
    dev = torch.device("cuda", 0)
    CFG = bench.CFG
    B, N, U = 1024, 20, CFG["users"]
    model = bench.build_model(pkg, U, dev, 0)
    opt = optim.FusedAdam(model.parameters(), CFG["lr"], CFG["betas"], CFG["eps"], 0)
    opt.set_skip_flags(model.error_flags())
    g = torch.Generator(device=dev).manual_seed(1)
    u = torch.randint(0, U, (B,), generator=g, device=dev)
    pos = torch.randn(B, 128, CFG["frames"], generator=g, device=dev)
    neg = torch.randn(B, N, 128, CFG["frames"], generator=g, device=dev)
    gstep = pkg.GraphedTrainStep(model, CFG["margin"], u, pos, neg, warmup=3)

    def step():
        loss = gstep()
        opt.step()
        return loss.detach().clone()

    def run(n, tag):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        host = []
        evs[0].record()
        for i in range(n):
            t0 = time.perf_counter()
            step()
            host.append((time.perf_counter() - t0) * 1e3)
            evs[i + 1].record()
        torch.cuda.synchronize()
        devt = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
        print(tag, "device ms:", " ".join("%.2f" % x for x in devt))
        print(tag, "host   ms:", " ".join("%.2f" % x for x in host), flush=True)

    run(6, "after capture      ")
    run(6, "after synchronize  ")
    if len(sys.argv) > 1:
        p = subprocess.Popen([sys.executable, "-c", "import time; time.sleep(0.5)"])
        run(6, "after fork         ")
        p.wait()
    torch.cuda.synchronize()
    time.sleep(0.3)
    run(6, "after 0.3 s idle   ")


if __name__ == "__main__":
    main()
