"""GPU: a discriminator update, compute_loss and a BC-mixed PPO update fed by the device-resident DeviceExpertLoader
(uint8 table in HBM, fused uint8 gather) against the same steps fed by the reference-style host fp32 batches of the same
samples.  The kernels see bit-identical normalised images either way (test_kernels_gpu.py::test_gather_mixup_metrics);
the update itself is only reproducible up to the order of its fp32 / fp64 atomic reductions (bias column sums, gradient
norms), so scalars are compared at rtol 1e-5 and parameters with the Adam-step bound of the parity tests
(max |diff| <= 2.5*lr per step, mean |diff| <= 0.05*lr)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

HP = dict(lr=1e-4, eps=1e-8, betas=(0.9, 0.99), clip_param=0.1, value_loss_coef=0.5, max_grad_norm=0.5,
          gail_lr=2.5e-4, gail_eps=1e-8, gail_betas=(0.9, 0.99), gail_max_grad_norm=0.5, logstd=[-1.4, -3.2])


class _HostLoader:
    def __init__(self, dev_loader):
        self.inner, self.batch_size = dev_loader, dev_loader.batch_size
    def __len__(self):
        return len(self.inner)
    def __iter__(self):
        for b in self.inner:
            yield tuple(t.cpu() for t in b)


def test_resident_expert_data_equals_host_batches_on_gpu():
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    from gail_carla_b200.expert import DeviceExpertLoader, ExpertDataset
    ds = ExpertDataset(os.path.join(GOLDEN, "expert_ds"), routes=[0, 3], n_eps=1)
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    results = []
    for resident in (True, False):
        torch.manual_seed(3)
        pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False).to("cuda")
        agent = G.PPO(pol, HP["clip_param"], 1, 4, HP["value_loss_coef"], "cuda", lr=HP["lr"], eps=HP["eps"], betas=HP["betas"],
                      max_grad_norm=HP["max_grad_norm"], gamma=0.3, decay=0.9, act_space=asp)
        disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, "cuda", HP["gail_lr"], HP["gail_eps"], HP["gail_betas"],
                               HP["gail_max_grad_norm"]).to("cuda")
        ro = G.RolloutStorage(4, 2, synthetic.OBS_SHAPE, (4,), (2,), device="cuda")
        synthetic.fill_rollout(ro, seed=13)
        dev_loader = DeviceExpertLoader(ds, batch_size=4, device="cuda")
        loader = dev_loader if resident else _HostLoader(dev_loader)
        torch.manual_seed(21)
        d_out = disc.update(loader, ro)
        loss = disc.compute_loss(loader, ro)
        ro.compute_returns(0.99, 0.95)
        p_out = agent.update(ro, loader)
        results.append(([float(v) for v in d_out], [float(v) for v in loss], [float(v) for v in p_out],
                        {k: v.detach().cpu().clone() for k, v in disc.state_dict().items()},
                        {k: v.detach().cpu().clone() for k, v in pol.state_dict().items()}))
    a, b = results
    for i in range(3):
        np.testing.assert_allclose(np.asarray(a[i]), np.asarray(b[i]), rtol=1e-5, atol=1e-7)
    for sd_a, sd_b, lr, steps in ((a[3], b[3], HP["gail_lr"], 2), (a[4], b[4], HP["lr"], 2)):
        for k in sd_a:
            d = (sd_a[k].double() - sd_b[k].double()).abs()
            assert d.max().item() <= 2.5 * lr * steps and d.mean().item() <= 0.05 * lr, (k, d.max().item(), d.mean().item())
