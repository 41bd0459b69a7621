"""Drop-in for tools/storage.py ``RolloutStorage`` with the tensors resident in HBM.

Same constructor arguments, attributes (``obs, metrics, rewards, gail_rewards, value_preds, returns,
action_log_probs, actions, masks, num_steps, num_processes, step``) and methods as tools/storage.py:6-79; callers keep
indexing / assigning slices exactly as tools/learn.py:73-74,138,197,205 do.  ``compute_returns`` is one launch of the
GAE scan kernel instead of a Python loop over T; minibatches are addressed by index tensors so the gather happens
inside the consumer kernels.

``obs_dtype=torch.uint8`` (opt-in) keeps the observations as the bytes they are made of: the CARLA adapter builds every
observation as ``uint8 / 255`` (carla_env.py:134-138), so a byte store is lossless, 4x smaller in HBM (29 -> 7.3 GB at
64 envs x 1024 steps) and 4x cheaper to upload and to gather.  ``rollouts.obs`` is then a :class:`ByteObs` tensor
(uint8, byte b stands for the fp32 value b/255 - exactly what ``ToTensor`` produces); assigning floating-point
observations (``insert``, ``rollouts.obs[0].copy_(obs)``, ``rollouts.obs[0] = obs``) quantises them and **raises** if any
value is not exactly k/255 - there is no silent rounding.  ``Policy`` / ``Discriminator`` accept uint8 observations.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch

from . import _abi as A


def quantize_obs_checked(obs: torch.Tensor) -> torch.Tensor:
    """fp32 observations on the uint8/255 grid -> uint8; raises ValueError when a value is off the grid."""
    q = torch.round(obs.detach().float() * 255.0).clamp_(0, 255).to(torch.uint8)
    if not bool((q.float() / 255.0 == obs.float()).all()):
        raise ValueError("uint8 observation store: observations must be exactly k/255 (carla_env.py:134-138 builds them "
                         "that way); use the default fp32 store for arbitrary floating-point observations")
    return q


class ByteObs(torch.Tensor):
    """uint8 observation tensor whose bytes stand for b/255.  Writing floating-point data into it goes through
    :func:`quantize_obs_checked` instead of torch's truncating float->uint8 cast."""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func in (torch.Tensor.copy_, torch.Tensor.__setitem__) and len(args) >= 2:
            k = 1 if func is torch.Tensor.copy_ else 2
            src = args[k] if len(args) > k else None
            if isinstance(src, torch.Tensor) and src.is_floating_point() and args[0].dtype == torch.uint8:
                q = quantize_obs_checked(src).to(args[0].device)
                args = args[:k] + (q,) + args[k + 1:]
        return super().__torch_function__(func, types, args, kwargs)

    def as_float(self) -> torch.Tensor:
        """The fp32 observations these bytes stand for (materialised)."""
        return self.as_subclass(torch.Tensor).float().div_(255.0)


class RolloutStorage(object):
    def __init__(self, num_steps, num_processes, obs_shape, metrics_shape, action_shape, device=None, obs_dtype=torch.float32):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=self.device)
        if obs_dtype == torch.uint8:
            self.obs = torch.zeros(num_steps + 1, num_processes, *obs_shape, dtype=torch.uint8, device=self.device).as_subclass(ByteObs)
        elif obs_dtype == torch.float32:
            self.obs = z(num_steps + 1, num_processes, *obs_shape)
        else:
            raise ValueError("obs_dtype must be torch.float32 (reference layout) or torch.uint8 (byte store)")
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
        # multi-GPU "exact" sharding: this storage holds envs [rank*N, (rank+1)*N) of a global
        # rollout of world*N envs; None = stand-alone storage (each rank permutes its own shard)
        self.shard: Optional[Tuple[int, int]] = None

    # tools/storage.py:21-30
    def insert(self, obs, metrics, actions, action_log_probs, value_preds, rewards, masks):
        s = self.step
        self.obs[s + 1].copy_(obs, non_blocking=True)      # ByteObs: exactness-checked quantisation of fp32 input
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
    def _to_device(self, idx: torch.Tensor) -> torch.Tensor:
        if self.device.type == "cuda":
            return idx.pin_memory().to(self.device, non_blocking=True)
        return idx

    def minibatch_indices(self, mini_batch_size: int, batch_size: Optional[int] = None) -> Iterator[torch.Tensor]:
        """tools/storage.py:57-63: BatchSampler(SubsetRandomSampler(range(n)), mb, drop_last=True).  The permutation
        is drawn from torch's default CPU generator exactly like SubsetRandomSampler does, then moved to the device;
        flat index = t * num_processes + n (time-major, tools/storage.py:66)."""
        if batch_size is None:
            batch_size = self.num_processes * self.num_steps
        perm = self._to_device(torch.randperm(batch_size))
        for s in range(0, batch_size - mini_batch_size + 1, mini_batch_size):
            yield perm[s:s + mini_batch_size]

    def set_shard(self, rank: int, world: int) -> None:
        """Declare this storage to be env shard `rank` of `world` equal shards of one global rollout (exact mode)."""
        self.shard = (int(rank), int(world)) if world > 1 else None

    def sharded_minibatches(self, global_mini_batch_size: int, batch_size: Optional[int] = None):
        """Exact multi-GPU sharding of tools/storage.py:57-66: ONE permutation of the global flat index space
        ``t * (world*N) + n_global`` is drawn from the default CPU generator (every rank draws the same one from the same
        seed - and the same one a single process holding all world*N envs would draw), cut into global minibatches, and
        each rank keeps the members whose env it owns.  Yields ``(pos, idx)``: ``pos`` (CPU int64) = positions inside the
        global minibatch owned by this rank, ``idx`` (device int64) = the corresponding local flat indices t*N+n."""
        rank, world = self.shard if self.shard is not None else (0, 1)
        N = self.num_processes
        Ng = N * world
        if batch_size is None:
            batch_size = Ng * self.num_steps
        perm = torch.randperm(batch_size)
        for s in range(0, batch_size - global_mini_batch_size + 1, global_mini_batch_size):
            g = perm[s:s + global_mini_batch_size]
            n_glob = g % Ng
            mine = (n_glob // N) == rank
            pos = torch.nonzero(mine).view(-1)
            local = (g[pos] // Ng) * N + (n_glob[pos] - rank * N)
            yield pos, self._to_device(local.contiguous())

    def flat(self, name: str) -> torch.Tensor:
        """[T(+1), N, ...] -> [(T(+1))*N, ...] view; rows t*N+n with t < T are the minibatch-addressable samples."""
        t = getattr(self, name)
        if isinstance(t, ByteObs):
            t = t.as_subclass(torch.Tensor)
        return t.view(t.shape[0] * t.shape[1], *t.shape[2:])

    # tools/storage.py:52-79 - API-compatible generator (device tensors).  The fused update paths do not use it.
    def feed_forward_generator(self, advantages, mini_batch_size, batch_size=None, only_last_cycle=False):
        for idx in self.minibatch_indices(mini_batch_size, batch_size):
            obs = self.flat("obs")[idx]
            if obs.dtype == torch.uint8:
                obs = obs.float().div_(255.0)
            yield (obs, self.flat("metrics")[idx], self.flat("actions")[idx],
                   self.flat("value_preds")[idx], self.flat("returns")[idx], self.flat("masks")[idx],
                   self.flat("action_log_probs")[idx], None if advantages is None else advantages.view(-1, 1)[idx])
