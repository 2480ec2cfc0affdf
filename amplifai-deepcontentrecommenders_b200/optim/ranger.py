"""Ranger = Rectified Adam (Liu et al. 2019) + Lookahead (Zhang et al. 2019), API-compatible with
the reference's dcrecommend/optim/ranger.py (same constructor, hyper-parameters, state keys
``step / exp_avg / exp_avg_sq / slow_buffer`` and update rule, ranger.py:82-165).

CUDA parameters take ONE multi-tensor kernel launch per step (csrc/optim.cu ``dcue_ranger_multi_step``: moments,
weight decay, rectified update and the lookahead interpolation fused, 1 read-modify-write pass over p/m/v(/slow)
instead of ~12 elementwise passes per parameter).  Parameters that are not on a CUDA device (host-side unit tests of the
update rule) take the same arithmetic as ``torch._foreach`` updates.
"""
import math

import torch
from torch.optim.optimizer import Optimizer

from ._multi_tensor import PointerTable


class Ranger(Optimizer):

    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999), eps=1e-5,
                 weight_decay=0):
        if not 0.0 <= alpha <= 1.0:
            raise ValueError(f"Invalid slow update rate: {alpha}")
        if not 1 <= k:
            raise ValueError(f"Invalid lookahead steps: {k}")
        if not lr > 0:
            raise ValueError(f"Invalid Learning Rate: {lr}")
        if not eps > 0:
            raise ValueError(f"Invalid eps: {eps}")
        defaults = dict(lr=lr, alpha=alpha, k=k, step_counter=0, betas=betas, N_sma_threshhold=N_sma_threshhold,
                        eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.N_sma_threshhold = N_sma_threshhold
        self.alpha = alpha
        self.k = k
        self._tables = {}
        self.skip_flags = []

    def set_skip_flags(self, flags):
        """See FusedAdam.set_skip_flags."""
        flags = [f for f in flags if f is not None]
        if len(flags) > 2:
            raise ValueError("at most two skip flags")
        self.skip_flags = flags

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}

    def add_param_group(self, param_group):
        super().add_param_group(param_group)
        self._tables = {}

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}
        self.__dict__.setdefault("skip_flags", [])
        for name in ("N_sma_threshhold", "alpha", "k"):
            if name not in self.__dict__:
                setattr(self, name, self.defaults[name])

    @staticmethod
    def _rectification(step, beta1, beta2, threshold):
        """-> (use_adaptive, step_size) of RAdam at `step` (variance rectification term)."""
        beta2_t = beta2 ** step
        n_max = 2.0 / (1.0 - beta2) - 1.0
        n_sma = n_max - 2.0 * step * beta2_t / (1.0 - beta2_t)
        bias1 = 1.0 - beta1 ** step
        if n_sma > threshold:
            r = math.sqrt((1.0 - beta2_t) * (n_sma - 4.0) / (n_max - 4.0) * (n_sma - 2.0) / n_sma * n_max / (n_max - 2.0))
            return True, r / bias1
        return False, 1.0 / bias1

    def _fused(self, key, group, step, ps):
        from .. import _lib as L
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("Ranger (fused): parameters must be contiguous fp32 CUDA tensors")
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                p.grad = p.grad.float().contiguous()
        beta1, beta2 = group["betas"]
        adaptive, step_size = self._rectification(step, beta1, beta2, self.N_sma_threshhold)
        tab = self._tables.setdefault(key, PointerTable())
        table, blk_first, n, blocks = tab.get(
            ps, [(self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"], self.state[p]["slow_buffer"]) for p in ps])
        fl = [f.data_ptr() for f in self.skip_flags] + [None, None]
        L.call("dcue_ranger_multi_step", table.data_ptr(), blk_first.data_ptr(), n, blocks, float(step_size * group["lr"]),
               float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"] * group["lr"]), int(adaptive),
               int(step % group["k"] == 0), float(self.alpha), fl[0], fl[1], L.stream())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if any(p.is_cuda for g in self.param_groups for p in g["params"][:1]):
            from .. import ops
            ops.join_backward_side()
        for gi, group in enumerate(self.param_groups):
            beta1, beta2 = group["betas"]
            # parameters of one group may be at different steps (e.g. late-added): bucket by step
            buckets = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Ranger optimizer does not support sparse gradients")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, dtype=torch.float32)
                    state["exp_avg_sq"] = torch.zeros_like(p, dtype=torch.float32)
                    state["slow_buffer"] = p.detach().clone()
                state["step"] += 1
                buckets.setdefault((state["step"], p.is_cuda), []).append(p)
            for (step, on_gpu), ps in buckets.items():
                if on_gpu:
                    self._fused((gi, step % 2 if len(buckets) > 1 else 0), group, step, ps)
                    continue
                grads = [p.grad.float() for p in ps]
                m = [self.state[p]["exp_avg"] for p in ps]
                v = [self.state[p]["exp_avg_sq"] for p in ps]
                torch._foreach_mul_(v, beta2)
                torch._foreach_addcmul_(v, grads, grads, value=1.0 - beta2)
                torch._foreach_mul_(m, beta1)
                torch._foreach_add_(m, grads, alpha=1.0 - beta1)
                adaptive, step_size = self._rectification(step, beta1, beta2, self.N_sma_threshhold)
                if group["weight_decay"] != 0:
                    torch._foreach_mul_(ps, 1.0 - group["weight_decay"] * group["lr"])
                if adaptive:
                    denom = torch._foreach_sqrt(v)
                    torch._foreach_add_(denom, group["eps"])
                    torch._foreach_addcdiv_(ps, m, denom, value=-step_size * group["lr"])
                else:
                    torch._foreach_add_(ps, m, alpha=-step_size * group["lr"])
                if step % group["k"] == 0:  # lookahead: slow += alpha * (fast - slow); fast = slow
                    slow = [self.state[p]["slow_buffer"] for p in ps]
                    torch._foreach_lerp_(slow, ps, self.alpha)
                    torch._foreach_copy_(ps, slow)
        return loss
