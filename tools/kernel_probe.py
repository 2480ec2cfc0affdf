"""Launch the step's big kernels at cfg2 size a few times (for `ncu --set full -k regex:...` captures and quick timings).
  python tools/kernel_probe.py [S]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 21504
res = bench.kernel_roofline(pkg, S, torch.device("cuda"))
print(json.dumps({k: {"ms": round(v["ms"], 4), "achieved": round(v["achieved"], 1)} for k, v in res.items()}))
