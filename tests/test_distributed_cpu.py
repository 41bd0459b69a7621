"""world_size-2 gloo test of the multi-rank host logic (env-sharded storage, advantage-statistics all-reduce,
gradient all-reduce before the fused clip+Adam) with the kernels replaced by their CPU statements."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace as NS
    from gail_carla_b200 import _abi
    from oracle import abi_emu
    for name in dir(abi_emu):
        fn = getattr(abi_emu, name)
        if callable(fn) and not name.startswith("_") and hasattr(_abi, name) and name not in ("call", "load_library"):
            setattr(_abi, name, fn)
    _abi.EMULATED = True
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    torch.set_num_threads(2)
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    torch.manual_seed(1)
    pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, [-1.4, -3.2], False)
    agent = G.PPO(pol, 0.1, 1, 4, 0.5, "cpu", lr=1e-4, eps=1e-8, betas=(0.9, 0.99), max_grad_norm=0.5)
    T, N = 4, 1
    ro = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device="cpu")
    synthetic.fill_rollout(ro, seed=50 + rank)
    ro.returns[:-1] = ro.value_preds[:-1] + torch.randn(T, N, 1, generator=torch.Generator().manual_seed(70 + rank))
    # global advantage statistics == statistics of the concatenated shards
    stats = torch.zeros(4, dtype=torch.float64)
    _abi.adv_stats(ro.returns, ro.value_preds, stats, T * N)
    dist.all_reduce(stats[:3])
    adv = (ro.returns[:-1] - ro.value_preds[:-1]).double().view(-1)
    gathered = [torch.zeros_like(adv) for _ in range(world)]
    dist.all_gather(gathered, adv)
    alladv = torch.cat(gathered)
    ok_stats = bool(torch.allclose(stats[0], alladv.sum()) and torch.allclose(stats[1], (alladv ** 2).sum()) and stats[2] == alladv.numel())
    torch.manual_seed(200 + rank)
    out = agent.update(ro)
    flat = pol.engine.flat.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    q.put((rank, ok_stats, bool(torch.equal(flat, ref)), float(out[0]), float(flat.abs().sum())))
    dist.destroy_process_group()


def test_two_rank_update_keeps_replicas_identical():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_stats, same, vloss, norm in res:
        assert ok_stats, f"rank {rank}: all-reduced advantage statistics differ from the concatenated shards"
        assert same, f"rank {rank}: parameters diverged from rank 0 after the all-reduced update"
        assert norm > 0
