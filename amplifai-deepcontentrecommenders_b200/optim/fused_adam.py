"""Adam as ONE multi-tensor CUDA launch per step (csrc/optim.cu), a drop-in for the ``torch.optim.Adam`` the reference
builds in ``DCUE._init_nn`` (dcrecommend/nn/dcue.py:143-147): same constructor arguments, update rule and
``state_dict`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so optimizer checkpoints interchange.
The dense [U, 300] user-table gradient makes the optimizer an HBM pass over four table-sized arrays every step."""
from __future__ import annotations

import torch
from torch.optim.optimizer import Optimizer

from .. import _lib as L


class FusedAdam(Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if not lr >= 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not eps >= 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}     # group index -> (key, table, blk_first, n_tensors, total_blocks)

    def _table(self, gi, params):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) for p in params)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit
        per = L.lib().dcue_adam_elems_per_block()
        rows, first, blocks = [], [], 0
        for p in params:
            st = self.state[p]
            rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()])
            first.append(blocks)
            blocks += (p.numel() + per - 1) // per
        dev = params[0].device
        table = torch.tensor(rows, dtype=torch.int64).to(dev)
        blk_first = torch.tensor(first, dtype=torch.int32).to(dev)
        hit = (key, table, blk_first, len(params), blocks)
        self._tables[gi] = hit
        return hit

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters with dense gradients "
                                       "(the DCUE B200 path has no CPU fallback)")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
            step = float(self.state[params[0]]["step"])
            if any(float(self.state[p]["step"]) != step for p in params):
                raise RuntimeError("FusedAdam: parameters of one group must share the step count")
            b1, b2 = group["betas"]
            _, table, blk_first, n, blocks = self._table(gi, params)
            L.call("dcue_adam_multi_step", table.data_ptr(), blk_first.data_ptr(), n, blocks, float(group["lr"]), float(b1), float(b2),
                   float(group["eps"]), float(group["weight_decay"]), 1.0 - b1 ** step, 1.0 - b2 ** step, L.stream())
        return loss
