"""Golden vectors for the expert data path.  Run in the build container only:

    python tests/golden/make_expert_golden.py

Writes a tiny data set in the reference's on-disk format (tests/golden/expert_ds: 2 routes x 1 episode, 5 + 4 steps,
192x192 RGB PNGs made of rectangles so they compress to ~1 KB), loads it with the UNMODIFIED reference classes
(`algo.wdgail.ExpertDataset` + `torch.utils.data.DataLoader(shuffle=True, drop_last=True)` as wdail_carla.py:161-183 does)
and stores what they produce: every item's metrics / actions / image checksum and a strided image sample, and the batch
composition of two epochs drawn after torch.manual_seed(7) (identified by the metrics rows)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from gail_carla_b200.expert import write_episode  # noqa: E402


def make_images(n, g):
    obs = torch.zeros(n, 3, 192, 192, dtype=torch.uint8)
    for i in range(n):
        for c, levels in enumerate(((255,), (255,), (120, 255))):
            for _ in range(3):
                y0, x0 = (int(v) for v in torch.randint(0, 150, (2,), generator=g))
                h, w = (int(v) for v in torch.randint(8, 42, (2,), generator=g))
                obs[i, c, y0:y0 + h, x0:x0 + w] = levels[int(torch.randint(0, len(levels), (1,), generator=g))]
    return obs


def main():
    ds_dir = os.path.join(HERE, "expert_ds")
    g = torch.Generator().manual_seed(11)
    for route, n in ((0, 5), (3, 4)):
        obs = make_images(n, g)
        metrics = torch.cat([torch.randn(n, 2, generator=g) * 1e-3, torch.rand(n, 1, generator=g) * 8,
                             torch.randint(1, 7, (n, 1), generator=g).float()], 1)
        actions = torch.cat([torch.randn(n, 1, generator=g) * 0.1, torch.rand(n, 1, generator=g)], 1)
        write_episode(ds_dir, route, 0, obs, metrics.tolist(), actions.tolist())

    from algo.wdgail import ExpertDataset as RefDataset   # the unmodified reference
    ref = RefDataset(ds_dir, routes=[0, 3], n_eps=1)
    out = {"length": np.int64(len(ref))}
    items = [ref[j] for j in range(len(ref))]
    out["metrics"] = torch.stack([it[1] for it in items]).numpy()
    out["actions"] = torch.stack([it[2] for it in items]).numpy()
    obs = torch.stack([it[0] for it in items])
    out["obs_sum"] = obs.double().sum((1, 2, 3)).numpy()
    out["obs_sample"] = obs[:, :, ::16, ::16].numpy()
    torch.manual_seed(7)
    loader = torch.utils.data.DataLoader(ref, batch_size=4, shuffle=True, drop_last=True)
    epochs = []
    for _ in range(2):
        epochs.append(np.stack([b[1].numpy() for b in loader]))       # metrics rows identify the samples
    out["epoch_metrics"] = np.stack(epochs)
    out["rand_after"] = torch.rand(1).numpy()                          # default-generator state after two epochs
    np.savez_compressed(os.path.join(HERE, "expert.npz"), **out)
    print("wrote", ds_dir, "and expert.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
