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


def _exact_worker(rank, world, port, q):
    """Exact-mode sharding: the 2-rank run on env shards must reproduce the UNMODIFIED reference's single-process
    result on the concatenated envs (tests/golden/update_tiny2.npz) - same global permutations, same mix-up draws,
    global-batch gradients before the clip (tools/storage.py:60-66, algo/ppo.py:115-119, algo/wdgail.py:112-145)."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace as NS
    import numpy as np
    from gail_carla_b200 import _abi
    from oracle import abi_emu
    for name in dir(abi_emu):
        fn = getattr(abi_emu, name)
        if callable(fn) and not name.startswith("_") and hasattr(_abi, name) and name not in ("call", "load_library"):
            setattr(_abi, name, fn)
    _abi.EMULATED = True
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    from gail_carla_b200.driver import update_iteration
    import test_host_cpu as H
    torch.set_num_threads(2)
    z = np.load(os.path.join(ROOT, "tests", "golden", "update_tiny2.npz"))
    T, N, B_ppo, B_gail, ppo_epoch, gail_epoch, n_expert, bc, seed = (int(v) for v in z["config"][:9])
    Nl = N // world
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    torch.manual_seed(seed)
    pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, H.HP["logstd"], False)
    agent = G.PPO(pol, H.HP["clip_param"], ppo_epoch, B_ppo // world, H.HP["value_loss_coef"], "cpu", lr=H.HP["lr"], eps=H.HP["eps"],
                  betas=H.HP["betas"], max_grad_norm=H.HP["max_grad_norm"], gamma=None, decay=None, act_space=asp)
    disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, "cpu", H.HP["gail_lr"], H.HP["gail_eps"], H.HP["gail_betas"],
                           H.HP["gail_max_grad_norm"])
    agent.exact_sharding = disc.exact_sharding = True
    full = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device="cpu")
    synthetic.fill_rollout(full, seed=seed + 10)
    ro = G.RolloutStorage(T, Nl, synthetic.OBS_SHAPE, (4,), (2,), device="cpu")
    for k in ("obs", "metrics", "actions", "action_log_probs", "value_preds", "returns", "masks", "gail_rewards", "rewards"):
        getattr(ro, k).copy_(getattr(full, k)[:, rank * Nl:(rank + 1) * Nl])
    ro.set_shard(rank, world)
    loader = synthetic.SyntheticExpertLoader(n_expert, B_gail, seed=seed + 20)     # the same global batches on every rank
    torch.manual_seed(seed + 100)
    d_out, p_out, cl0, cl1 = update_iteration(pol, agent, disc, ro, loader, gamma=H.HP["gamma"], gae_lambda=H.HP["gae_lambda"],
                                              gail_epoch=gail_epoch, bcgail=False, diagnostics=True)
    err = None
    try:
        def close(a, b, what, rtol=2e-3, atol=2e-4):
            a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
            assert np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True), f"{what}: {a} vs {b}"
        close(cl0, z["compute_loss_before"], "compute_loss before")
        close(d_out, z["disc_update"], "Discriminator.update tuples")
        close(cl1, z["compute_loss_after"], "compute_loss after")
        close(ro.gail_rewards.numpy(), z["gail_rewards"][:, rank * Nl:(rank + 1) * Nl], "gail_rewards shard")
        close(ro.returns.numpy(), z["returns"][:, rank * Nl:(rank + 1) * Nl], "returns shard")
        close([np.nan if x is None else x for x in p_out], z["ppo_update"], "PPO.update tuple")
        H.digest_check(disc.state_dict(), z, "disc", H.HP["gail_lr"], gail_epoch * min(n_expert, T * N // B_gail))
        H.digest_check(pol.state_dict(), z, "pol", H.HP["lr"], ppo_epoch * (T * N // B_ppo))
    except AssertionError as ex:
        err = str(ex)[:500]
    q.put((rank, err))
    dist.destroy_process_group()


def test_two_rank_exact_sharding_matches_single_process_reference():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + os.getpid() % 1000
    procs = [ctx.Process(target=_exact_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err in res:
        assert err is None, f"rank {rank}: {err}"


def _empty_share_worker(rank, world, port, q):
    """Exact sharding when a rank owns NO row of a global minibatch: it must contribute a zero gradient through the same
    sequence of collectives (same buckets, same order) as the ranks that ran a backward pass - a mismatch dead-locks NCCL
    (seen at 8 ranks x 4 rows) and corrupts gloo.  Checked against a one-process replay of the concatenated envs."""
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
    from gail_carla_b200 import synthetic, optim
    from gail_carla_b200.driver import update_iteration
    import test_host_cpu as H
    torch.set_num_threads(2)
    T, Nl, Bglob = 3, 1, world     # one row per rank on average: empty shares are frequent
    N = Nl * world
    sp, asp = NS(shape=(4,)), NS(shape=(2,))

    def build(mb):
        torch.manual_seed(5)
        pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, H.HP["logstd"], False)
        agent = G.PPO(pol, H.HP["clip_param"], 2, mb, H.HP["value_loss_coef"], "cpu", lr=H.HP["lr"], eps=H.HP["eps"], betas=H.HP["betas"],
                      max_grad_norm=H.HP["max_grad_norm"], gamma=None, decay=None, act_space=asp)
        disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, "cpu", H.HP["gail_lr"], H.HP["gail_eps"], H.HP["gail_betas"],
                               H.HP["gail_max_grad_norm"])
        return pol, agent, disc

    full = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device="cpu")
    synthetic.fill_rollout(full, seed=17)
    loader = synthetic.SyntheticExpertLoader(3, Bglob, seed=23)
    keys = ("obs", "metrics", "actions", "action_log_probs", "value_preds", "returns", "masks", "gail_rewards", "rewards")
    # how many (rank, minibatch) shares are empty under the seed used below?  (same draws as the sharded run will make)
    torch.manual_seed(99)
    empties = 0
    probe = G.RolloutStorage(T, Nl, (1, 2, 2), (4,), (2,), device="cpu"); probe.set_shard(rank, world)
    for _ in range(3):                                   # compute_loss-free run: critic epoch, 2 PPO epochs
        empties += sum(int(pos.numel() == 0) for pos, _ in probe.sharded_minibatches(Bglob))
    pol, agent, disc = build(Bglob // world)
    agent.exact_sharding = disc.exact_sharding = True
    ro = G.RolloutStorage(T, Nl, synthetic.OBS_SHAPE, (4,), (2,), device="cpu")
    for k in keys:
        getattr(ro, k).copy_(getattr(full, k)[:, rank * Nl:(rank + 1) * Nl])
    ro.set_shard(rank, world)
    torch.manual_seed(99)
    d_out, p_out = update_iteration(pol, agent, disc, ro, loader, gamma=0.99, gae_lambda=0.95, gail_epoch=1)
    with optim.single_process():
        pol1, agent1, disc1 = build(Bglob)
        torch.manual_seed(99)
        d1, p1 = update_iteration(pol1, agent1, disc1, full, loader, gamma=0.99, gae_lambda=0.95, gail_epoch=1)
    err = None
    try:
        for (k, a), (_, b) in zip(list(pol.state_dict().items()) + list(disc.state_dict().items()),
                                  list(pol1.state_dict().items()) + list(disc1.state_dict().items())):
            # fp32 sums in a different order (per-rank partial gradients summed by the all-reduce vs one batch) move an Adam
            # step by a small fraction of lr (1e-4 / 2.5e-4): 2e-5 absolute
            assert torch.allclose(a, b, rtol=1e-4, atol=2e-5), f"{k}: max abs diff {(a - b).abs().max().item():.3g}"
        assert torch.allclose(torch.tensor(d_out[0]), torch.tensor(d1[0]), rtol=1e-4, atol=1e-6)
        assert torch.allclose(torch.tensor([x for x in p_out if x is not None]), torch.tensor([x for x in p1 if x is not None]),
                              rtol=1e-4, atol=1e-6)
    except AssertionError as ex:
        err = str(ex)[:500]
    q.put((rank, err, empties))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_exact_sharding_with_empty_rank_shares(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + 7 * world + os.getpid() % 1000
    procs = [ctx.Process(target=_empty_share_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(e for _, _, e in res) > 0, "the seed no longer produces an empty share - pick another one"
    for rank, err, _ in res:
        assert err is None, f"rank {rank}: {err}"
