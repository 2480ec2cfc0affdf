"""CUDA-graph capture of the fixed-shape part of a DCUE training step (forward + fused score/hinge loss +
backward: ~100 kernel launches) so that a step costs one graph launch instead of ~100 launch gaps.
The optimizer / scheduler stay outside the graph (the reference's Python-float learning-rate schedule keeps
working unchanged).  Under data parallelism (`dp=DataParallelDCUE(model)`) the NCCL collectives -- the per-layer
BatchNorm statistic all-reduces and the flat gradient all-reduce -- are captured into the same graph, so a
multi-GPU step is one graph launch as well.
"""
from __future__ import annotations

import os

import torch

from . import ops


class GraphedTrainStep:
    """step = GraphedTrainStep(model, margin, u, pos, neg)   # example batch fixes the shapes
       loss = step(u, pos, neg); optimizer.step()            # gradients are in model.parameters()[i].grad

    With `pool` given, (pos, neg) are int64 song-index tensors into the resident pool (index feed)."""

    def __init__(self, model, margin, u, pos, neg, pool=None, warmup=3, dp=None):
        self.model, self.margin, self.pool, self.dp = model, margin, pool, dp
        # Under data parallelism the host is kept at most `max_inflight` graph launches ahead of the device: with an
        # unbounded launch queue the ranks measured 0.4 ms per step slower (8 x B200: graph launches carrying NCCL work).
        self.max_inflight = int(os.environ.get("DCUE_DP_MAX_INFLIGHT", "3")) if dp is not None else 0
        self._events, self._launches = [], 0
        self.u, self.pos, self.neg = u.clone(), pos.clone(), neg.clone()
        # warm-up and capture run training-mode forwards: BatchNorm running statistics / num_batches_tracked would advance
        # by warmup + 1 steps behind the caller's back -> snapshot them here and restore after the capture
        buffers = [(b, b.detach().clone()) for b in model.buffers()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # warm-up off the default stream: allocators, workspaces, cub temp
            for _ in range(warmup):
                model.zero_grad(set_to_none=True)
                self._loss().backward()
                if dp is not None:
                    dp.reduce_gradients()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        params = [p for p in model.parameters() if p.requires_grad]
        self._params = params
        # Gradient storage.  Default ("owned"): the capture starts with every .grad = None, so autograd ASSIGNS the tensors
        # the backward kernels produce (graph-pool memory: the same addresses on every replay) instead of accumulating into
        # pre-existing ones -- the round-2 timeline showed 29 ATen add kernels + the zero fill per step (~75 us of 2.6 ms)
        # doing nothing but that.  __call__ re-binds p.grad to this graph's tensors, so several graphs over one model (dense
        # and index feed) and eager steps in between (zero_grad(set_to_none=False)) keep working.
        # DCUE_GRAPH_GRADS=accumulate: the graph zeroes the EXISTING .grad tensors and accumulates into them in place.
        self.owned = os.environ.get("DCUE_GRAPH_GRADS", "owned") != "accumulate"
        if not self.owned:
            for p in params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            self._grads = [p.grad for p in params]
            # one more eager pass on the STATIC gradient tensors: everything keyed on their addresses (the peer gradient
            # all-reduce's pointer table, ...) is built now -- a capture must not contain pageable host-to-device copies
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                torch._foreach_zero_(self._grads)
                self._loss().backward()
                if dp is not None:
                    dp.reduce_gradients()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        self.passes = warmup + (1 if self.owned else 2)      # forward+backward passes run by this constructor (launch accounting)
        self.graph = torch.cuda.CUDAGraph()
        if self.owned:
            for p in params:
                p.grad = None
        with torch.cuda.graph(self.graph):
            if not self.owned:
                torch._foreach_zero_(self._grads)
            self.loss = self._loss()
            self.loss.backward()
            ops.join_backward_side()        # every side stream of the backward pass re-joins before the capture ends
            if dp is not None:
                dp.reduce_gradients()
        if self.owned:
            self._grads = [p.grad for p in params]
        with torch.no_grad():
            for b, saved in buffers:
                b.copy_(saved)
        torch.cuda.synchronize()

    def release(self):
        """Drop the captured graph (and the NCCL work it holds).  Call on every rank before
        torch.distributed.destroy_process_group(): a live graph with captured collectives keeps the communicator busy."""
        torch.cuda.synchronize()
        self.graph = None
        self.loss = None

    def _loss(self):
        if self.dp is not None:
            if self.pool is not None:
                return self.dp.loss_step_indexed(self.u, self.pool, self.pos, self.neg, self.margin)
            return self.dp.loss_step(self.u, self.pos, self.neg, self.margin)
        if self.pool is not None:
            return self.model.hinge_loss_step_indexed(self.u, self.pool, self.pos, self.neg, self.margin)
        return self.model.hinge_loss_step(self.u, self.pos, self.neg, self.margin)

    def __call__(self, u=None, pos=None, neg=None):
        """Replay on a new batch of the captured shapes (pass nothing to reuse the static inputs).
        Returns the loss tensor (static: read it before the next call)."""
        if u is not None:
            self.u.copy_(u, non_blocking=True)
            self.pos.copy_(pos, non_blocking=True)
            self.neg.copy_(neg, non_blocking=True)
        if self.max_inflight > 0:
            if len(self._events) < self.max_inflight:
                self._events.append(torch.cuda.Event())
            ev = self._events[self._launches % self.max_inflight]
            if self._launches >= self.max_inflight:
                ev.synchronize()          # the launch issued max_inflight steps ago has finished
            self.graph.replay()
            ev.record()
            self._launches += 1
        else:
            self.graph.replay()
        if self.owned:
            for p, g in zip(self._params, self._grads):
                if p.grad is not g:
                    p.grad = g
        return self.loss
