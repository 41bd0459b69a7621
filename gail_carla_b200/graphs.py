"""CUDA-graph replay of a fixed-shape minibatch step.

One PPO / critic optimisation step is ~70-130 kernel launches whose arguments do not change from minibatch to minibatch
once the inputs have been gathered into the workspace: same buffers, same shapes, same tensor maps.  At the per-rank
minibatch sizes of an 8-GPU run (512 rows) the launches are shorter than the host needs to issue them, so the step is
captured once into a CUDA graph (kernels, memsets, the NCCL all-reduce of the gradient buckets and the side stream they
run on) and replayed with ONE launch per minibatch.  Everything that does change between replays lives in device
memory the graph reads: the minibatch's row indices, the mix-up coefficients and Adam's step-dependent scalars.

``StepGraph.run(key, fn)`` executes ``fn`` exactly once per call: eagerly the first time a key is seen (buffers get
allocated, cuDNN-style lazy state settles), captured-then-replayed the second time, replayed afterwards.  A failed capture
(or a CPU / emulated device) falls back to eager execution for good - the result is the same either way.
"""
from __future__ import annotations

from typing import Callable, Hashable, Optional

import torch

from . import _abi as A

import weakref

ENABLED = True        # module switch (bench.py --no-graphs, the instrumented per-kernel timing pass)
_ALL = weakref.WeakSet()


def release_all() -> None:
    """Drop every captured graph.  Call this (after a device synchronize) BEFORE ``torch.distributed.destroy_process_group()``:
    graphs that captured NCCL collectives keep communicator resources alive, and tearing the process group down underneath
    them blocked ProcessGroupNCCL's watchdog until it aborted the process (seen on B200, torch 2.11 / NCCL 2.28)."""
    for g in list(_ALL):
        g.reset()


class _Entry:
    __slots__ = ("graph", "seen", "launches")

    def __init__(self):
        self.graph, self.seen, self.launches = None, 0, 0


class StepGraph:
    MAX_ENTRIES = 4       # e.g. the two observation buffers of a double-buffered rollout upload

    def __init__(self, name: str):
        self.name = name
        self.entries = {}
        self.broken = False
        _ALL.add(self)

    def reset(self) -> None:
        self.entries = {}

    def run(self, key: Hashable, fn: Callable[[], None], device) -> None:
        if not ENABLED or self.broken or torch.device(device).type != "cuda" or getattr(A, "EMULATED", False):
            fn()
            return
        e = self.entries.get(key)
        if e is None:
            if len(self.entries) >= self.MAX_ENTRIES:
                self.entries.pop(next(iter(self.entries)))
            e = self.entries[key] = _Entry()
        if e.graph is not None:
            e.graph.replay()
            A.LAUNCHES += e.launches
            return
        e.seen += 1
        if e.seen == 1:                         # first step with these buffers / shapes: eager (allocates every lazy buffer)
            fn()
            return
        before = A.LAUNCHES
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
        except Exception as ex:                 # capture not possible here: run eagerly from now on
            self.broken = True
            self.entries = {}
            import warnings
            warnings.warn(f"gail_carla_b200: CUDA-graph capture of the {self.name} step failed ({type(ex).__name__}: "
                          f"{str(ex)[:200]}); continuing with eager launches")
            torch.cuda.synchronize()
            fn()
            return
        e.launches = A.LAUNCHES - before        # launches recorded during capture did not execute yet
        A.LAUNCHES = before
        e.graph = g
        g.replay()
        A.LAUNCHES += e.launches
