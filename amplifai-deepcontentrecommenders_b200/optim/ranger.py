"""Ranger = Rectified Adam (Liu et al. 2019) + Lookahead (Zhang et al. 2019), API-compatible with
the reference's dcrecommend/optim/ranger.py (same constructor, hyper-parameters, state keys
``step / exp_avg / exp_avg_sq / slow_buffer`` and update rule), written as multi-tensor
(``torch._foreach``) updates: a handful of fused launches per step instead of ~12 per parameter.
"""
import math

import torch
from torch.optim.optimizer import Optimizer


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

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        for group in self.param_groups:
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
                buckets.setdefault(state["step"], []).append(p)
            for step, ps in buckets.items():
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
