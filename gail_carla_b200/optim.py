"""``clip_grad_norm_`` + ``optim.Adam`` (algo/ppo.py:43,115-119; algo/wdgail.py:35,140-145) as two launches over the
flat parameter / gradient buffers, with the NCCL gradient all-reduce in between when a process group is active.

The object is a real ``torch.optim.Optimizer`` so the reference's drivers can keep mutating
``optimizer.param_groups[i]['lr']`` (tools/utli.py:121-125, tools/learn.py:102-106).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import _abi as A
from .engine import FlatParams


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, flat_getter, params, lr, eps, betas, max_grad_norm: Optional[float]):
        super().__init__(list(params), dict(lr=lr, eps=eps, betas=tuple(betas)))
        self._flat_getter = flat_getter
        self.max_grad_norm = max_grad_norm
        self.t = 0
        self._m = self._v = self._sumsq = None
        self._owner = None

    def _buffers(self, flat: FlatParams):
        if self._m is None or self._owner is not flat.flat or self._m.device != flat.flat.device:
            old_m, old_v = self._m, self._v
            self._m = torch.zeros_like(flat.flat)
            self._v = torch.zeros_like(flat.flat)
            if old_m is not None and old_m.numel() == self._m.numel():   # parameters were moved: keep the moments
                self._m.copy_(old_m); self._v.copy_(old_v)
            self._sumsq = torch.zeros(1, dtype=torch.float64, device=flat.flat.device)
            self._owner = flat.flat

    @torch.no_grad()
    def step(self, closure=None):
        flat: FlatParams = self._flat_getter()
        self._buffers(flat)
        g = self.param_groups[0]
        n = flat.numel
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            # replicas hold equal-size minibatch shards: mean of the per-rank mean-gradients == global-batch gradient
            dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
            flat.grad.mul_(1.0 / dist.get_world_size())
        self.t += 1
        b1, b2 = g["betas"]
        if self.max_grad_norm is not None:
            self._sumsq.zero_()
            A.grad_sumsq(flat.grad, n, self._sumsq)
        A.clip_adam(flat.flat, flat.grad, self._m, self._v, n, self._sumsq, self.max_grad_norm, float(g["lr"]), float(b1),
                    float(b2), float(g["eps"]), 1.0 - b1 ** self.t, 1.0 - b2 ** self.t)

    def zero_grad(self, set_to_none: bool = False):   # gradients live in the flat buffer; never detach them
        flat: FlatParams = self._flat_getter()
        flat.grad.zero_()
