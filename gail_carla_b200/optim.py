"""``clip_grad_norm_`` + ``optim.Adam`` (algo/ppo.py:43,115-119; algo/wdgail.py:35,140-145) as two launches over the
flat parameter / gradient buffers, with the NCCL gradient all-reduce in between when a process group is active.

The object is a real ``torch.optim.Optimizer`` so the reference's drivers can keep mutating
``optimizer.param_groups[i]['lr']`` (tools/utli.py:121-125, tools/learn.py:102-106).

Multi-GPU: every rank holds the gradient of its rows of the global minibatch.  The engines hand
finished slices of the flat gradient buffer to :class:`GradReducer` while the backward pass is still running (the fully
connected layers - 90 % of the bytes - are final before the convolution gradients start), so the NCCL all-reduce runs
on a side stream under the remaining dgrad / wgrad launches; ``step()`` reduces whatever is left, joins the streams and
runs norm + clip + Adam.  The 1/world factor of the gradient mean and ``optimizer.zero_grad()`` are folded into those
two kernels (no extra passes over the gradient).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _abi as A
from .engine import FlatParams


_FORCE_SINGLE = False


def world_size() -> int:
    if _FORCE_SINGLE:
        return 1
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class single_process:
    """Context manager: run the update classes as a stand-alone process even though a process group is initialised
    (used to replay a sharded run on the concatenated envs inside one rank for the multi-GPU parity check)."""

    def __enter__(self):
        global _FORCE_SINGLE
        self._old, _FORCE_SINGLE = _FORCE_SINGLE, True
        return self

    def __exit__(self, *exc):
        global _FORCE_SINGLE
        _FORCE_SINGLE = self._old
        return False


class GradReducer:
    """Sum-all-reduces slices of a flat gradient buffer on a side stream, in the order they become final."""

    def __init__(self):
        self.stream: Optional[torch.cuda.Stream] = None
        self.done: List[Tuple[int, int]] = []

    def _side(self, dev) -> Optional[torch.cuda.Stream]:
        if dev.type != "cuda":
            return None
        if self.stream is None or self.stream.device != dev:
            self.stream = torch.cuda.Stream(device=dev)
        return self.stream

    def ready(self, flat: FlatParams, lo: int, hi: int) -> None:
        """Elements [lo, hi) of ``flat.grad`` are final on the current stream: start reducing them."""
        if world_size() == 1 or hi <= lo:
            return
        side = self._side(flat.grad.device)
        if side is None:                       # gloo / CPU tests: reduce in line
            dist.all_reduce(flat.grad[lo:hi], op=dist.ReduceOp.SUM)
        else:
            side.wait_stream(torch.cuda.current_stream(flat.grad.device))
            with torch.cuda.stream(side):
                dist.all_reduce(flat.grad[lo:hi], op=dist.ReduceOp.SUM)
        self.done.append((lo, hi))

    def finish(self, flat: FlatParams) -> None:
        """Reduce every element not handed over through ``ready`` and make the current stream wait for all of it."""
        if world_size() == 1:
            self.done = []
            return
        pos = 0
        for lo, hi in sorted(self.done) + [(flat.numel, flat.numel)]:
            if lo > pos:
                self.ready(flat, pos, lo)
            pos = max(pos, hi)
        self.done = []
        if self.stream is not None and flat.grad.device.type == "cuda":
            torch.cuda.current_stream(flat.grad.device).wait_stream(self.stream)


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, flat_getter, params, lr, eps, betas, max_grad_norm: Optional[float]):
        super().__init__(list(params), dict(lr=lr, eps=eps, betas=tuple(betas)))
        self._flat_getter = flat_getter
        self.max_grad_norm = max_grad_norm
        self.t = 0
        self._m = self._v = self._sumsq = None
        self._owner = None
        self.reducer = GradReducer()
        # gradient scale applied inside the norm / Adam kernels: None = 1/world (mean of per-rank mean gradients, equal
        # shards); exact mode sets 1.0 (per-rank gradients are already weighted by 1/B_global and only need summing)
        self.grad_scale: Optional[float] = None
        self._hyper_host = self._hyper_dev = None

    def _buffers(self, flat: FlatParams):
        if self._m is None or self._owner is not flat.flat or self._m.device != flat.flat.device:
            old_m, old_v = self._m, self._v
            self._m = torch.zeros_like(flat.flat)
            self._v = torch.zeros_like(flat.flat)
            if old_m is not None and old_m.numel() == self._m.numel():   # parameters were moved: keep the moments
                self._m.copy_(old_m); self._v.copy_(old_v)
            self._sumsq = torch.zeros(1, dtype=torch.float64, device=flat.flat.device)
            self._owner = flat.flat
            self._hyper_host = self._hyper_dev = None

    # ---- step-dependent scalars kept on the device so a captured graph of the step can be replayed -----------
    def device_hyper(self, flat: FlatParams) -> torch.Tensor:
        if self._hyper_dev is None:
            dev = flat.flat.device
            self._hyper_dev = torch.zeros(3, device=dev)
        return self._hyper_dev

    def begin_schedule(self, n_steps: int) -> None:
        """Upload {lr, 1-beta1^t, 1-beta2^t} for the next `n_steps` optimiser steps in one copy (the learning rate of
        ``param_groups`` is read here, i.e. once per update like the reference's per-update schedule,
        tools/learn.py:102-106).  ``advance()`` then selects the row of the step about to run with a device-to-device
        copy, so neither an eager step nor a graph replay needs a host-side scalar."""
        flat: FlatParams = self._flat_getter()
        self._buffers(flat)
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        dev = flat.flat.device
        host = torch.empty(max(n_steps, 1), 3, pin_memory=dev.type == "cuda")
        for i in range(n_steps):
            t = self.t + i + 1
            host[i, 0] = float(g["lr"]); host[i, 1] = 1.0 - b1 ** t; host[i, 2] = 1.0 - b2 ** t
        self._table_host = host                                  # keep the pinned source alive until the next schedule
        self._table = host.to(dev, non_blocking=True)
        self._table_i = 0

    def advance(self) -> None:
        """t += 1; stage this step's scalars where ``step(from_device_hyper=True)`` reads them."""
        flat: FlatParams = self._flat_getter()
        self.t += 1
        self.device_hyper(flat).copy_(self._table[self._table_i], non_blocking=True)
        self._table_i += 1

    @torch.no_grad()
    def step(self, closure=None, *, from_device_hyper: bool = False):
        flat: FlatParams = self._flat_getter()
        self._buffers(flat)
        g = self.param_groups[0]
        n = flat.numel
        world = world_size()
        self.reducer.finish(flat)
        scale = self.grad_scale if self.grad_scale is not None else 1.0 / world
        if not from_device_hyper:
            self.t += 1
        b1, b2 = g["betas"]
        if self.max_grad_norm is not None:
            A.grad_sumsq(flat.grad, n, self._sumsq, scale)      # (clears the accumulator on the stream first)
        A.clip_adam(flat.flat, flat.grad, self._m, self._v, n, self._sumsq, self.max_grad_norm, float(g["lr"]), float(b1),
                    float(b2), float(g["eps"]), 1.0 - b1 ** max(self.t, 1), 1.0 - b2 ** max(self.t, 1), scale, True,
                    self.device_hyper(flat) if from_device_hyper else None)
        flat.grad_clean = True

    def zero_grad(self, set_to_none: bool = False):   # gradients live in the flat buffer; never detach them
        flat: FlatParams = self._flat_getter()
        flat.grad.zero_()
        flat.grad_clean = True
