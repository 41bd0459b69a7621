"""RolloutStorage host semantics against restatements of tools/storage.py lines (CPU tensors; no kernels involved):
insert / step wrap-around (:21-30), after_update (:32-35) and feed_forward_generator (:52-79), whose minibatches must
be the ones BatchSampler(SubsetRandomSampler(range(n)), mb, drop_last=True) draws from the same seed."""
import torch
from torch.utils.data.sampler import BatchSampler, SubsetRandomSampler


def _filled(T=6, N=3, seed=0):
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    ro = G.RolloutStorage(T, N, (3, 8, 8), (4,), (2,), device="cpu")
    g = torch.Generator().manual_seed(seed)
    for name in ("obs", "metrics", "actions", "value_preds", "returns", "action_log_probs", "masks", "gail_rewards"):
        t = getattr(ro, name)
        t.copy_(torch.randn(t.shape, generator=g))
    return ro


def test_insert_and_after_update_follow_the_reference():
    import gail_carla_b200 as G
    T, N = 3, 2
    ro = G.RolloutStorage(T, N, (3, 8, 8), (4,), (2,), device="cpu")
    assert ro.obs.shape == (T + 1, N, 3, 8, 8) and ro.masks.shape == (T + 1, N, 1) and bool((ro.masks == 1).all())
    assert ro.rewards.shape == ro.gail_rewards.shape == ro.action_log_probs.shape == (T, N, 1) and ro.actions.shape == (T, N, 2)
    g = torch.Generator().manual_seed(1)
    steps = []
    for k in range(T + 1):                    # one more than T: the step index wraps (tools/storage.py:30)
        item = [torch.randn(N, 3, 8, 8, generator=g), torch.randn(N, 4, generator=g), torch.randn(N, 2, generator=g),
                torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g),
                (torch.rand(N, 1, generator=g) > 0.5).float()]
        s = ro.step
        ro.insert(*item)
        steps.append((s, item))
        assert ro.step == (s + 1) % T
    s, item = steps[-1]                       # the wrapped insert overwrote slot 0 / 1
    assert s == 0
    assert torch.equal(ro.obs[1], item[0]) and torch.equal(ro.metrics[1], item[1]) and torch.equal(ro.actions[0], item[2])
    assert torch.equal(ro.action_log_probs[0], item[3]) and torch.equal(ro.value_preds[0], item[4])
    assert torch.equal(ro.rewards[0], item[5]) and torch.equal(ro.masks[1], item[6])
    last = (ro.obs[-1].clone(), ro.metrics[-1].clone(), ro.masks[-1].clone())
    ro.after_update()
    assert torch.equal(ro.obs[0], last[0]) and torch.equal(ro.metrics[0], last[1]) and torch.equal(ro.masks[0], last[2])


def test_feed_forward_generator_draws_the_reference_minibatches():
    ro = _filled()
    T, N = ro.num_steps, ro.num_processes
    adv = torch.randn(T, N, 1, generator=torch.Generator().manual_seed(9))
    for batch_size in (None, 12):
        torch.manual_seed(123)
        got = list(ro.feed_forward_generator(adv, 4, batch_size=batch_size))
        torch.manual_seed(123)
        n = T * N if batch_size is None else batch_size
        sampler = BatchSampler(SubsetRandomSampler(range(n)), 4, drop_last=True)     # tools/storage.py:60-63
        ref = []
        for indices in sampler:                                                         # tools/storage.py:65-79
            ref.append((ro.obs[:-1].view(-1, *ro.obs.size()[2:])[indices], ro.metrics[:-1].view(-1, 4)[indices],
                        ro.actions.view(-1, 2)[indices], ro.value_preds[:-1].view(-1, 1)[indices],
                        ro.returns[:-1].view(-1, 1)[indices], ro.masks[:-1].view(-1, 1)[indices],
                        ro.action_log_probs.view(-1, 1)[indices], adv.view(-1, 1)[indices]))
        assert len(got) == len(ref) == n // 4
        for a, b in zip(got, ref):
            assert len(a) == 8
            for x, y in zip(a, b):
                assert torch.equal(x, y)
    torch.manual_seed(5)
    first = next(iter(ro.feed_forward_generator(None, 4)))
    assert first[7] is None
