"""Drop-in for tools/storage.py ``RolloutStorage`` with the tensors resident in HBM.

Same constructor arguments, attributes (``obs, metrics, rewards, gail_rewards, value_preds, returns,
action_log_probs, actions, masks, num_steps, num_processes, step``) and methods as tools/storage.py:6-79; callers keep
indexing / assigning slices exactly as tools/learn.py:73-74,138,197,205 do.  ``compute_returns`` is one launch of the
GAE scan kernel instead of a Python loop over T; minibatches are addressed by index tensors so the gather happens
inside the consumer kernels.
"""
from __future__ import annotations

from typing import Iterator, Optional

import torch

from . import _abi as A


class RolloutStorage(object):
    def __init__(self, num_steps, num_processes, obs_shape, metrics_shape, action_shape, device=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        self.obs = z(num_steps + 1, num_processes, *obs_shape)
        self.metrics = z(num_steps + 1, num_processes, *metrics_shape)
        self.rewards = z(num_steps, num_processes, 1)
        self.gail_rewards = z(num_steps, num_processes, 1)
        self.value_preds = z(num_steps + 1, num_processes, 1)
        self.returns = z(num_steps + 1, num_processes, 1)
        self.action_log_probs = z(num_steps, num_processes, 1)
        self.actions = z(num_steps, num_processes, *action_shape)
        self.masks = torch.ones(num_steps + 1, num_processes, 1, dtype=torch.float32, device=self.device)
        self.num_steps = num_steps
        self.num_processes = num_processes
        self.step = 0
        # {sum, sum of squares, count} of returns - value_preds, refreshed by compute_returns (algo/ppo.py:47-49)
        self.adv_stats = torch.zeros(4, dtype=torch.float64, device=self.device)

    # tools/storage.py:21-30
    def insert(self, obs, metrics, actions, action_log_probs, value_preds, rewards, masks):
        s = self.step
        self.obs[s + 1].copy_(obs, non_blocking=True)
        self.metrics[s + 1].copy_(metrics, non_blocking=True)
        self.actions[s].copy_(actions, non_blocking=True)
        self.action_log_probs[s].copy_(action_log_probs, non_blocking=True)
        self.value_preds[s].copy_(value_preds, non_blocking=True)
        self.rewards[s].copy_(rewards, non_blocking=True)
        self.masks[s + 1].copy_(masks, non_blocking=True)
        self.step = (s + 1) % self.num_steps

    # tools/storage.py:32-35
    def after_update(self):
        self.obs[0].copy_(self.obs[-1])
        self.metrics[0].copy_(self.metrics[-1])
        self.masks[0].copy_(self.masks[-1])

    # tools/storage.py:37-50 (gail_coef = 1, env_coef = 0): one segmented reverse scan over time per env
    def compute_returns(self, gamma, gae_lambda):
        A.gae_returns(self.gail_rewards, self.value_preds, self.masks, self.returns, float(gamma), float(gae_lambda),
                      None, self.adv_stats)

    # ---- minibatch addressing -------------------------------------------------------------------------------
    def minibatch_indices(self, mini_batch_size: int, batch_size: Optional[int] = None) -> Iterator[torch.Tensor]:
        """tools/storage.py:57-63: BatchSampler(SubsetRandomSampler(range(n)), mb, drop_last=True).  The permutation
        is drawn from torch's default CPU generator exactly like SubsetRandomSampler does, then moved to the device;
        flat index = t * num_processes + n (time-major, tools/storage.py:66)."""
        if batch_size is None:
            batch_size = self.num_processes * self.num_steps
        perm = torch.randperm(batch_size)
        if self.device.type == "cuda":
            perm = perm.pin_memory().to(self.device, non_blocking=True)
        for s in range(0, batch_size - mini_batch_size + 1, mini_batch_size):
            yield perm[s:s + mini_batch_size]

    def flat(self, name: str) -> torch.Tensor:
        """[T(+1), N, ...] -> [(T(+1))*N, ...] view; rows t*N+n with t < T are the minibatch-addressable samples."""
        t = getattr(self, name)
        return t.view(t.shape[0] * t.shape[1], *t.shape[2:])

    # tools/storage.py:52-79 - API-compatible generator (device tensors).  The fused update paths do not use it.
    def feed_forward_generator(self, advantages, mini_batch_size, batch_size=None, only_last_cycle=False):
        for idx in self.minibatch_indices(mini_batch_size, batch_size):
            yield (self.flat("obs")[idx], self.flat("metrics")[idx], self.flat("actions")[idx],
                   self.flat("value_preds")[idx], self.flat("returns")[idx], self.flat("masks")[idx],
                   self.flat("action_log_probs")[idx], None if advantages is None else advantages.view(-1, 1)[idx])
