"""Drop-in for common/running_mean_std.py.

``mean`` / ``var`` / ``count`` are float64 like the reference (common/running_mean_std.py:5-8).  ``update`` accepts a
numpy array (host path, identical arithmetic) or a 1-D CUDA tensor; for CUDA tensors the batch moments and the Chan
merge run on the device (gc_welford_merge) and the state stays in HBM until one of the attributes is read.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi as A


def update_mean_var_count_from_moments(mean, var, count, batch_mean, batch_var, batch_count):
    """common/running_mean_std.py:20-31 (Chan et al. parallel variance merge)."""
    delta = batch_mean - mean
    tot = count + batch_count
    m2 = var * count + batch_var * batch_count + np.square(delta) * count * batch_count / tot
    return mean + delta * batch_count / tot, m2 / tot, tot


class RunningMeanStd(object):
    def __init__(self, epsilon=1e-4, shape=()):
        self._mean = np.zeros(shape, "float64")
        self._var = np.ones(shape, "float64")
        self._count = epsilon
        self._dev_state = None      # double[3] = {mean, var, count} on the device while device updates are pending

    def _pull(self):
        if self._dev_state is not None:
            m, v, c = self._dev_state.cpu().tolist()
            self._mean, self._var, self._count = np.array(m, "float64"), np.array(v, "float64"), c
            self._dev_state = None

    @property
    def mean(self):
        self._pull()
        return self._mean

    @mean.setter
    def mean(self, value):
        self._pull()
        self._mean = value

    @property
    def var(self):
        self._pull()
        return self._var

    @var.setter
    def var(self, value):
        self._pull()
        self._var = value

    @property
    def count(self):
        self._pull()
        return self._count

    @count.setter
    def count(self, value):
        self._pull()
        self._count = value

    def update(self, x):
        if isinstance(x, torch.Tensor) and x.is_cuda and self._mean.shape == () and x.dim() == 1:
            if self._dev_state is None:
                self._dev_state = torch.tensor([float(self._mean), float(self._var), float(self._count)],
                                               dtype=torch.float64, device=x.device)
                self._scratch = torch.zeros(2, dtype=torch.float64, device=x.device)
            A.welford_merge(self._dev_state, x.float().contiguous(), self._scratch)
            return
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        self._pull()
        self._mean, self._var, self._count = update_mean_var_count_from_moments(
            self._mean, self._var, self._count, batch_mean, batch_var, batch_count)
