"""Expert demonstrations: the reference's on-disk format, held once in HBM.

Reference: ``ExpertDataset`` (algo/wdgail.py:192-241) + ``DataLoader(..., batch_size, shuffle=True, drop_last=True)``
(wdail_carla.py:161-183); writer carla_exp.py:38-80.  Layout on disk::

    <dataset_directory>/route_%02d/ep_%02d/episode.json           pandas JSON, columns 'actions' [2], 'metrics' [4]
    <dataset_directory>/route_%02d/ep_%02d/birdview_masks/%04d_00.png   192x192 RGB uint8

The reference decodes a PNG the first time a sample is drawn, keeps the fp32 tensor (442 KB) on the host and re-uploads
1.8 GB per 4096-batch on every discriminator step.  Here the data set is decoded once into a **uint8 table**
``[L,3,192,192]`` (110 KB per sample - the PNGs' own precision), and ``DeviceExpertLoader`` keeps that table in HBM:
a batch is just an index vector; the image gather kernel reads the bytes, applies ToTensor's ``/255``, the
normalisation and the space-to-depth transform in one pass (``gc_gather_obs_u8_s2d``), so a discriminator step moves no
expert bytes over PCIe and reads 4x fewer from HBM.  Shuffling reproduces ``DataLoader(shuffle=True)``'s sampler draw
for draw (same consumption of the default generator), so runs are comparable sample for sample.
"""
from __future__ import annotations

import json
import os
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch


def _read_episode(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """episode.json written by ``DataFrame({'actions': [...], 'metrics': [...]}).to_json`` (carla_exp.py:73-78):
    ``{column: {row index as string: value}}``; rows are taken in index order like ``route_df.iloc[i]``."""
    with open(path) as fh:
        d = json.load(fh)
    keys = sorted(d["actions"].keys(), key=int)
    actions = np.asarray([d["actions"][k] for k in keys], dtype=np.float32).reshape(len(keys), -1)
    metrics = np.asarray([d["metrics"][k] for k in keys], dtype=np.float32).reshape(len(keys), -1)
    return actions, metrics


def _read_mask(path: str) -> torch.Tensor:
    """PNG -> uint8 [3,H,W] (``Image.open(...).convert('RGB')`` + the CHW transpose of ``ToTensor``)."""
    from PIL import Image
    with Image.open(path) as im:
        a = np.asarray(im.convert("RGB"), dtype=np.uint8)
    return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))


class ExpertDataset(torch.utils.data.Dataset):
    """Drop-in for algo/wdgail.py:192-241 (same constructor, ``len`` and ``(obs, metrics, action)`` items with
    ``obs = uint8/255`` as fp32 ``[3,192,192]``); images are decoded eagerly into ``obs_u8``."""

    def __init__(self, dataset_directory, routes=1, n_eps=1, start=0):
        self.dataset_path = str(dataset_directory)
        if isinstance(routes, int):
            routes = [routes]
        self.get_idx: List[Tuple[int, int, int]] = []
        acts, mets, imgs = [], [], []
        for route_idx in routes:
            for ep_idx in range(start, start + n_eps):
                ep_dir = os.path.join(self.dataset_path, "route_%02d" % route_idx, "ep_%02d" % ep_idx)
                a, m = _read_episode(os.path.join(ep_dir, "episode.json"))
                for step_idx in range(a.shape[0]):
                    self.get_idx.append((route_idx, ep_idx, step_idx))
                    imgs.append(_read_mask(os.path.join(ep_dir, "birdview_masks", "{:0>4d}_{:0>2d}.png".format(step_idx, 0))))
                acts.append(torch.from_numpy(a))
                mets.append(torch.from_numpy(m))
        self.length = len(self.get_idx)
        self.trajs_actions = torch.cat(acts) if acts else torch.zeros(0, 2)
        self.trajs_metrics = torch.cat(mets) if mets else torch.zeros(0, 4)
        self.obs_u8 = torch.stack(imgs) if imgs else torch.zeros(0, 3, 192, 192, dtype=torch.uint8)

    @classmethod
    def from_tensors(cls, obs_u8: torch.Tensor, metrics: torch.Tensor, actions: torch.Tensor) -> "ExpertDataset":
        self = cls.__new__(cls)
        self.dataset_path = None
        self.length = int(obs_u8.shape[0])
        self.get_idx = [(0, 0, i) for i in range(self.length)]
        self.obs_u8, self.trajs_metrics, self.trajs_actions = obs_u8, metrics.float(), actions.float()
        return self

    def __len__(self):
        return self.length

    def __getitem__(self, j):
        return self.obs_u8[j].float().div(255), self.trajs_metrics[j], self.trajs_actions[j]


class DeviceBatch:
    """One expert batch of a device-resident data set: the tables plus the int64 row indices of the batch.  The
    discriminator / BC paths gather straight from the tables; iterating (``obs, metrics, action = batch``) materialises
    the reference's fp32 tensors for any other consumer."""

    __slots__ = ("obs_table", "metrics_table", "actions_table", "idx")

    def __init__(self, obs_table, metrics_table, actions_table, idx):
        self.obs_table, self.metrics_table, self.actions_table, self.idx = obs_table, metrics_table, actions_table, idx

    @property
    def batch_rows(self) -> int:
        return int(self.idx.shape[0])

    def __iter__(self):
        i = self.idx
        yield self.obs_table[i].float().div(255)
        yield self.metrics_table[i]
        yield self.actions_table[i]


class DeviceExpertLoader:
    """``DataLoader(ExpertDataset, batch_size, shuffle, drop_last)`` with the data set resident on `device`.

    Protocol of the hot path (algo/wdgail.py:101,112,158; algo/ppo.py:88-102): ``.batch_size``, ``len()``, iteration.
    Each epoch draws the permutation exactly as ``torch.utils.data`` does - one int64 from the default generator for the
    loader's base seed, one for ``RandomSampler``'s private generator, ``torch.randperm(n, generator=that)`` - so the
    batches contain the samples the reference's loader would produce from the same seed."""

    def __init__(self, dataset: ExpertDataset, batch_size: int, shuffle: bool = True, drop_last: bool = True, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.n = len(dataset)
        self.obs_table = dataset.obs_u8.to(self.device).contiguous()
        self.metrics_table = dataset.trajs_metrics.to(self.device, torch.float32).contiguous()
        self.actions_table = dataset.trajs_actions.to(self.device, torch.float32).contiguous()

    def __len__(self) -> int:
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def epoch_order(self) -> torch.Tensor:
        torch.empty((), dtype=torch.int64).random_()                       # _BaseDataLoaderIter: base seed
        if not self.shuffle:
            return torch.arange(self.n)
        seed = int(torch.empty((), dtype=torch.int64).random_().item())    # RandomSampler.__iter__
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(self.n, generator=g)

    def __iter__(self) -> Iterator[DeviceBatch]:
        order = self.epoch_order().to(self.device, non_blocking=True)      # one small upload per epoch
        stop = self.n - self.batch_size + 1 if self.drop_last else self.n
        for o in range(0, max(stop, 0), self.batch_size):
            yield DeviceBatch(self.obs_table, self.metrics_table, self.actions_table, order[o:o + self.batch_size].contiguous())


def expert_rows(batch, device):
    """(obs rows, metrics rows, action rows, idx or None, batch size) of an expert batch in either form: a
    ``DeviceBatch`` (tables + indices, nothing copied) or the reference's ``(obs, metrics, action)`` tensors (obs fp32, or
    uint8 bytes standing for b/255 - what ``ExpertDataset.obs_u8`` holds)."""
    if isinstance(batch, DeviceBatch):
        return batch.obs_table, batch.metrics_table, batch.actions_table, batch.idx, batch.batch_rows
    obs, met, act = batch
    if obs.dtype == torch.uint8:      # bytes standing for b/255 (ToTensor of the PNGs): gathered by gc_gather_obs_u8_s2d
        obs = obs.to(device, non_blocking=True).contiguous()
    else:
        obs = obs.to(device, torch.float32, non_blocking=True).contiguous()
    met = met.to(device, torch.float32, non_blocking=True).contiguous()
    act = act.to(device, torch.float32, non_blocking=True).contiguous()
    return obs, met, act, None, int(obs.shape[0])


def write_episode(dataset_directory, route_idx: int, ep_idx: int, obs_u8: torch.Tensor, metrics, actions) -> None:
    """Write one episode in the reference's format (carla_exp.py:38-80: episode.json + birdview_masks/%04d_00.png)."""
    from PIL import Image
    ep_dir = os.path.join(str(dataset_directory), "route_%02d" % route_idx, "ep_%02d" % ep_idx)
    os.makedirs(os.path.join(ep_dir, "birdview_masks"), exist_ok=True)
    n = int(obs_u8.shape[0])
    for i in range(n):
        Image.fromarray(obs_u8[i].permute(1, 2, 0).contiguous().numpy()).save(
            os.path.join(ep_dir, "birdview_masks", "{:0>4d}_{:0>2d}.png".format(i, 0)))
    doc = {"actions": {str(i): [float(v) for v in actions[i]] for i in range(n)},
           "metrics": {str(i): [float(v) for v in metrics[i]] for i in range(n)}}
    with open(os.path.join(ep_dir, "episode.json"), "w") as fh:
        json.dump(doc, fh)
