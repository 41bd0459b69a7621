"""Reference-API holes closed in round 2, on the emulated ABI (host logic, fp32): differentiable
``Policy.evaluate_actions`` (tools/model.py:45-53), ``Discriminator.forward(gp=True)`` (algo/wdgail.py:51-52), the
expert-loader / ``batch_size`` edge cases of ``Discriminator.update`` / ``compute_loss`` and the byte store."""
import pytest
import torch

import api_cases as AC


def test_evaluate_actions_backward_cpu(emulated_abi):
    # fp32 on both sides, but ONE pre-activation within ~1e-7 of zero whose LeakyReLU' mask flips under a different
    # summation order already moves the conv1/conv2 gradients by ~1e-3 relative (one element of 7e5 off by a factor 5);
    # the layers downstream of the flip agree to 1e-6.  Hence 5e-3 here, not the 2e-4 of tests/test_grads_cpu.py.
    AC.evaluate_actions_is_differentiable("cpu", 5, 0.9999, 5e-3)


def test_forward_gp_handles_cpu(emulated_abi):
    AC.forward_gp_returns_first_order_handles("cpu", 3, 2e-4)


def test_expert_loader_edge_cases_cpu(emulated_abi):
    AC.expert_loader_shorter_and_dropped_remainder("cpu", 2e-3)


def test_reward_inf_tail_cpu(emulated_abi):
    AC.reward_saturates_to_inf("cpu")


def test_byte_store_rejects_off_grid_values():
    import gail_carla_b200 as G
    ro = G.RolloutStorage(2, 1, (3, 192, 192), (4,), (2,), device="cpu", obs_dtype=torch.uint8)
    ok = torch.randint(0, 256, (1, 3, 192, 192)).float() / 255.0
    ro.obs[0].copy_(ok)
    assert torch.equal(ro.obs[0].as_float(), ok)
    with pytest.raises(ValueError):
        ro.obs[1].copy_(ok * 0.5 + 0.001)
    with pytest.raises(ValueError):
        ro.insert(ok + 1e-4, torch.zeros(1, 4), torch.zeros(1, 2), torch.zeros(1, 1), torch.zeros(1, 1), torch.zeros(1, 1),
                  torch.ones(1, 1))
    obs, *_ = next(ro.feed_forward_generator(None, 2))
    assert obs.dtype == torch.float32 and float(obs.max()) <= 1.0


def test_sharded_minibatches_partition_the_global_permutation():
    import gail_carla_b200 as G
    T, N, world, Bg = 5, 3, 2, 6
    torch.manual_seed(4)
    perm = torch.randperm(T * N * world)
    seen = []
    for rank in range(world):
        ro = G.RolloutStorage(T, N, (1, 2, 2), (4,), (2,), device="cpu")
        ro.set_shard(rank, world)
        torch.manual_seed(4)
        for b, (pos, idx) in enumerate(ro.sharded_minibatches(Bg)):
            g = perm[b * Bg:(b + 1) * Bg][pos]
            assert ((g % (N * world)) // N == rank).all()
            assert torch.equal(idx, (g // (N * world)) * N + (g % (N * world)) - rank * N)
            seen += [(b, int(p)) for p in pos]
    assert sorted(seen) == [(b, p) for b in range(T * N * world // Bg) for p in range(Bg)]
