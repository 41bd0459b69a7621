"""Expert data path (gail_carla_b200/expert.py) against what the UNMODIFIED reference classes produce on the committed
fixture (tests/golden/expert_ds + expert.npz, made by tests/golden/make_expert_golden.py): items bit-identical, the
loader's batch composition identical draw for draw with DataLoader(shuffle=True, drop_last=True), and the
discriminator / BC paths give the same result from a device-resident DeviceExpertLoader as from host fp32 batches."""
import os
from types import SimpleNamespace as NS

import numpy as np
import torch

from conftest import GOLDEN

HP = dict(lr=1e-4, eps=1e-8, betas=(0.9, 0.99), clip_param=0.1, value_loss_coef=0.5, max_grad_norm=0.5,
          gail_lr=2.5e-4, gail_eps=1e-8, gail_betas=(0.9, 0.99), gail_max_grad_norm=0.5, logstd=[-1.4, -3.2])


def _dataset():
    from gail_carla_b200.expert import ExpertDataset
    return ExpertDataset(os.path.join(GOLDEN, "expert_ds"), routes=[0, 3], n_eps=1)


def test_dataset_items_equal_the_reference():
    z = np.load(os.path.join(GOLDEN, "expert.npz"))
    ds = _dataset()
    assert len(ds) == int(z["length"]) == 9
    assert ds.get_idx[0] == (0, 0, 0) and ds.get_idx[5] == (3, 0, 0)
    items = [ds[j] for j in range(len(ds))]
    obs = torch.stack([it[0] for it in items])
    assert obs.dtype == torch.float32 and obs.shape == (9, 3, 192, 192)
    np.testing.assert_array_equal(torch.stack([it[1] for it in items]).numpy(), z["metrics"])
    np.testing.assert_array_equal(torch.stack([it[2] for it in items]).numpy(), z["actions"])
    np.testing.assert_array_equal(obs[:, :, ::16, ::16].numpy(), z["obs_sample"])
    np.testing.assert_array_equal(obs.double().sum((1, 2, 3)).numpy(), z["obs_sum"])


def test_loader_draws_the_batches_dataloader_would():
    from gail_carla_b200.expert import DeviceExpertLoader
    z = np.load(os.path.join(GOLDEN, "expert.npz"))
    loader = DeviceExpertLoader(_dataset(), batch_size=4, shuffle=True, drop_last=True, device="cpu")
    assert len(loader) == 2 and loader.batch_size == 4 and bool(loader)
    torch.manual_seed(7)
    for ep in range(2):
        batches = list(loader)
        assert len(batches) == 2
        for b, ref_metrics in zip(batches, z["epoch_metrics"][ep]):
            obs, met, act = b                       # generic consumers unpack the reference's tuple
            np.testing.assert_array_equal(met.numpy(), ref_metrics)
            assert obs.shape == (4, 3, 192, 192) and act.shape == (4, 2)
    np.testing.assert_array_equal(torch.rand(1).numpy(), z["rand_after"])   # same consumption of the default generator


def _models(seed):
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    torch.manual_seed(seed)
    pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False)
    agent = G.PPO(pol, HP["clip_param"], 1, 4, HP["value_loss_coef"], "cpu", lr=HP["lr"], eps=HP["eps"], betas=HP["betas"],
                  max_grad_norm=HP["max_grad_norm"], gamma=0.3, decay=0.9, act_space=asp)
    disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, "cpu", HP["gail_lr"], HP["gail_eps"], HP["gail_betas"],
                           HP["gail_max_grad_norm"])
    ro = G.RolloutStorage(4, 2, synthetic.OBS_SHAPE, (4,), (2,), device="cpu")
    synthetic.fill_rollout(ro, seed=seed + 10)
    return pol, agent, disc, ro


class _HostLoader:
    """The same batches as fp32 host tensors, i.e. what DataLoader(ExpertDataset) hands to the reference."""
    def __init__(self, dev_loader):
        self.inner, self.batch_size = dev_loader, dev_loader.batch_size
    def __len__(self):
        return len(self.inner)
    def __iter__(self):
        for b in self.inner:
            yield tuple(t.clone() for t in b)


def test_resident_batches_give_the_same_update_as_host_batches(emulated_abi):
    from gail_carla_b200.expert import DeviceExpertLoader
    results = []
    for resident in (True, False):
        pol, agent, disc, ro = _models(3)
        dev_loader = DeviceExpertLoader(_dataset(), batch_size=4, device="cpu")
        loader = dev_loader if resident else _HostLoader(dev_loader)
        torch.manual_seed(21)
        d_out = disc.update(loader, ro)
        loss = disc.compute_loss(loader, ro)
        ro.compute_returns(0.99, 0.95)
        p_out = agent.update(ro, loader)            # BC mix reads the first expert batch (algo/ppo.py:88-102)
        results.append((d_out, loss, p_out, {k: v.clone() for k, v in disc.state_dict().items()},
                        {k: v.clone() for k, v in pol.state_dict().items()}))
    a, b = results
    np.testing.assert_array_equal(np.asarray(a[0], dtype=np.float64), np.asarray(b[0], dtype=np.float64))
    np.testing.assert_array_equal(np.asarray(a[1], dtype=np.float64), np.asarray(b[1], dtype=np.float64))
    np.testing.assert_array_equal(np.asarray([float(v) for v in a[2]]), np.asarray([float(v) for v in b[2]]))
    for sd_a, sd_b in ((a[3], b[3]), (a[4], b[4])):
        for k in sd_a:
            assert torch.equal(sd_a[k], sd_b[k]), k
