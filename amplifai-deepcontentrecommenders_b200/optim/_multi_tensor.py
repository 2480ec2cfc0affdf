"""Device-side pointer tables for the multi-tensor optimizer kernels (csrc/optim.cu)."""
from __future__ import annotations

import torch

from .. import _lib as L


class PointerTable:
    """Rows of raw device pointers (+ element count) for every tensor of one param group, cached on the device.

    The cache key covers EVERY pointer stored in the table (parameter, gradient and all optimizer-state tensors): after
    ``load_state_dict`` replaces the state tensors, or ``zero_grad(set_to_none=True)`` re-allocates gradients, the table is
    rebuilt instead of pointing at freed memory.  Rebuilds go through a pinned staging buffer (one async copy)."""

    def __init__(self):
        self._hit = None
        self._stage = None

    def clear(self):
        self._hit = None

    def get(self, params, state_tensors):
        """state_tensors: list (per param) of tuples of state tensors, in kernel order -> (table, blk_first, n, blocks)."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), p.numel()) + tuple(t.data_ptr() for t in st)
                    for p, st in zip(params, state_tensors))
        if self._hit is not None and self._hit[0] == key:
            return self._hit[1:]
        per = L.lib().dcue_adam_elems_per_block()
        width = 3 + len(state_tensors[0])
        n = len(params)
        rows = torch.empty(n * width + n, dtype=torch.int64)
        blocks = 0
        for i, (p, st) in enumerate(zip(params, state_tensors)):
            rows[i * width: (i + 1) * width] = torch.tensor([p.data_ptr(), p.grad.data_ptr()] + [t.data_ptr() for t in st]
                                                            + [p.numel()], dtype=torch.int64)
            rows[n * width + i] = blocks
            blocks += (p.numel() + per - 1) // per
        if blocks >= 2 ** 31:
            raise RuntimeError("multi-tensor optimizer: too many elements for one launch")
        dev = params[0].device
        if self._stage is None or self._stage.numel() < rows.numel():
            self._stage = torch.empty(rows.numel(), dtype=torch.int64).pin_memory()
        else:
            torch.cuda.current_stream(dev).synchronize()      # the previous async copy out of the staging buffer
        self._stage[: rows.numel()].copy_(rows)
        d = self._stage[: rows.numel()].to(dev, non_blocking=True)
        table = d[: n * width]
        blk_first = d[n * width:].to(torch.int32)
        self._hit = (key, table, blk_first, n, blocks)
        return self._hit[1:]
