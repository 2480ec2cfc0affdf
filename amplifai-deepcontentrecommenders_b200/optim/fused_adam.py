"""Adam as ONE multi-tensor CUDA launch per step (csrc/optim.cu), a drop-in for the ``torch.optim.Adam`` the reference
builds in ``DCUE._init_nn`` (dcrecommend/nn/dcue.py:143-147): same constructor arguments, update rule and
``state_dict`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so optimizer checkpoints interchange.
The dense [U, 300] user-table gradient makes the optimizer an HBM pass over four table-sized arrays every step."""
from __future__ import annotations

import torch
from torch.optim.optimizer import Optimizer

from .. import _lib as L
from ._multi_tensor import PointerTable


class FusedAdam(Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if not lr >= 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not eps >= 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}     # group index -> PointerTable
        self.skip_flags = []  # up to two device int32 flags: a raised flag turns the step into a no-op (bad index in the batch)

    def set_skip_flags(self, flags):
        """Device error flags (e.g. DCUENet.error_flags()): when one is raised the kernel leaves every parameter and
        moment untouched, so a NaN-poisoned step never reaches the weights and no host sync is needed per step."""
        flags = [f for f in flags if f is not None]
        if len(flags) > 2:
            raise ValueError("at most two skip flags")
        self.skip_flags = flags

    def _invalidate(self):
        self._tables = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._invalidate()            # the state tensors were replaced: cached raw pointers are stale

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        if hasattr(self, "_tables"):
            self._invalidate()

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}
        self.__dict__.setdefault("skip_flags", [])

    def _table(self, gi, params):
        tab = self._tables.setdefault(gi, PointerTable())
        return tab.get(params, [(self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) for p in params])

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from .. import ops
        ops.join_backward_side()          # gradients of a backward side stream are complete before they are read
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
            table, blk_first, n, blocks = self._table(gi, params)
            fl = [f.data_ptr() for f in self.skip_flags] + [None, None]
            L.call("dcue_adam_multi_step", table.data_ptr(), blk_first.data_ptr(), n, blocks, float(group["lr"]), float(b1), float(b2),
                   float(group["eps"]), float(group["weight_decay"]), 1.0 - b1 ** step, 1.0 - b2 ** step, fl[0], fl[1], L.stream())
        return loss
