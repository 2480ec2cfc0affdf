"""Where a work item of the eval scorer spends its cycles (block 0, first appender warp): python tools/topk_phases.py users songs"""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
ev, L = pkg.eval, pkg._lib
nu, ni = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(3)
uf = torch.randn(nu, 100, generator=g, device="cuda")
itf = torch.randn(ni, 100, generator=g, device="cuda")
items = ev.normalize_factors(itf)
buf = (ctypes.c_ulonglong * 8)()
for mode in ("1", "0"):
    os.environ["DCUE_TOPK_2PASS"] = mode
    ev.topk_scores(uf[:4096], itf, 100, normalized_items=items)
    torch.cuda.synchronize()
    L.lib().dcue_topk_debug_cycles(ctypes.addressof(buf), 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ev.topk_scores(uf, itf, 100, normalized_items=items)
    e1.record()
    torch.cuda.synchronize()
    L.lib().dcue_topk_debug_cycles(ctypes.addressof(buf), 1)
    v = list(buf)
    print("2pass=%s  %.2f ms  items %d tiles %d | cycles: stream %d final-boundary %d select+sort %d tail %d" %
          (mode, e0.elapsed_time(e1), v[4], v[5], v[0], v[1], v[2], v[3]))
