"""Synthetic CARLA-shaped rollouts and expert batches.

Shapes and value ranges follow the simulator adapter the reference trains on:
obs ``[3,192,192]`` birdview masks on the uint8/255 grid (carla_env.py:134-138,
carla_gym/core/obs_manager/birdview/chauffeurnet.py:186-189), metrics
``[gps_x, gps_y, speed, command]`` (carla_env.py:144), action ``[steer, throttle]``
(carla_env.py:93-94).  Everything is drawn from an explicit ``torch.Generator`` so
the CPU oracle, the golden-vector script and the CUDA path see identical inputs.
"""
from __future__ import annotations

from typing import Iterator, Tuple

import torch

OBS_SHAPE = (3, 192, 192)
METRICS_SHAPE = (4,)
ACTION_SHAPE = (2,)


def _gen(seed: int, device="cpu") -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def synth_obs(n: int, g: torch.Generator, device="cpu") -> torch.Tensor:
    """[n,3,192,192] fp32 with ch0/ch1 in {0,1} and ch2 in {0,120/255,1}."""
    r = torch.randint(0, 6, (n, 3, 192, 192), generator=g, device=device, dtype=torch.uint8)
    lut = torch.tensor([[0, 0, 0, 255, 255, 255], [0, 0, 0, 0, 255, 255], [0, 0, 120, 120, 255, 255]],
                       dtype=torch.uint8, device=device)
    out = torch.empty((n, 3, 192, 192), dtype=torch.float32, device=device)
    for c in range(3):
        out[:, c] = lut[c][r[:, c].long()].float() / 255.0
    return out


def synth_metrics(n: int, g: torch.Generator, device="cpu") -> torch.Tensor:
    """[n,4]: x,y ~ N(0,1e-3^2) (GPS degrees), speed ~ U[0,8), command in {1..6}."""
    xy = torch.randn((n, 2), generator=g, device=device) * 1e-3
    v = torch.rand((n, 1), generator=g, device=device) * 8.0
    c = torch.randint(1, 7, (n, 1), generator=g, device=device).float()
    return torch.cat([xy, v, c], dim=1)


def synth_actions(n: int, g: torch.Generator, steer_std: float, device="cpu") -> torch.Tensor:
    steer = torch.randn((n, 1), generator=g, device=device) * steer_std
    thr = torch.rand((n, 1), generator=g, device=device)
    return torch.cat([steer, thr], dim=1)


def fill_rollout(ro, seed: int = 1, chunk: int = 256) -> None:
    """Fill a RolloutStorage-like object in place (obs, metrics, actions, log-probs, values, masks, gail_rewards).

    Works for storages on any device: draws happen on the storage's device in
    time-major chunks so the 29 GB / 58 GB configs never need a second copy.
    """
    dev = ro.obs.device
    g = _gen(seed, dev.type if dev.type == "cpu" else dev)
    T, N = ro.num_steps, ro.num_processes
    for t0 in range(0, T + 1, chunk):
        t1 = min(T + 1, t0 + chunk)
        ro.obs[t0:t1].copy_(synth_obs((t1 - t0) * N, g, dev).view(t1 - t0, N, *OBS_SHAPE))
    ro.metrics.copy_(synth_metrics((T + 1) * N, g, dev).view(T + 1, N, 4))
    ro.actions.copy_(synth_actions(T * N, g, 0.25, dev).view(T, N, 2))
    ro.action_log_probs.copy_(torch.randn((T, N, 1), generator=g, device=dev) * 0.5 + 1.0)
    ro.value_preds.copy_(torch.randn((T + 1, N, 1), generator=g, device=dev))
    m = (torch.rand((T + 1, N, 1), generator=g, device=dev) >= 1.0 / 400.0).float()
    m[0] = 1.0
    ro.masks.copy_(m)
    ro.gail_rewards.copy_(torch.nn.functional.softplus(torch.randn((T, N, 1), generator=g, device=dev)))
    ro.rewards.zero_()


class SyntheticExpertLoader:
    """Stands in for ``DataLoader(ExpertDataset, batch_size, shuffle, drop_last)`` (wdail_carla.py:161-183).

    Protocol used by the hot path (algo/wdgail.py:101,112,158; algo/ppo.py:88-102):
    ``.batch_size``, truthiness via ``__len__`` and iteration yielding CPU fp32
    ``(obs[B,3,192,192], metrics[B,4], action[B,2])``.  Batches are pre-drawn once (pinned
    when ``pin=True``) and replayed in order on every pass.
    """

    def __init__(self, n_batches: int, batch_size: int, seed: int = 2, pin: bool = False, obs_u8: bool = False):
        """obs_u8: yield the observations as the uint8 bytes they are made of (b stands for b/255, what the PNGs of
        algo/wdgail.py:222-227 hold) instead of fp32 - 4x fewer bytes to keep on the host and to upload."""
        self.batch_size = batch_size
        g = _gen(seed)
        self._batches = []
        for _ in range(n_batches):
            b = (synth_obs(batch_size, g), synth_metrics(batch_size, g), synth_actions(batch_size, g, 0.1))
            if obs_u8:
                b = (torch.round(b[0] * 255.0).to(torch.uint8),) + b[1:]
            if pin:
                b = tuple(t.pin_memory() for t in b)
            self._batches.append(b)

    def __len__(self) -> int:
        return len(self._batches)

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        return iter(self._batches)


class _Space:
    def __init__(self, shape):
        self.shape = tuple(shape)


class SyntheticVecEnv:
    """Simulator-free stand-in for the reference's vectorised CARLA env (tools/envs.py; carla_env.py:93-100,134-147):
    same protocol - ``reset() -> (obs[N,3,192,192], metrics[N,4])``, ``step(action[N,2]) -> (obs, metrics,
    rewards[N,1], done[N], infos[N])``, ``observation_space / metrics_space / action_space`` with ``.shape``;
    ``infos[i]`` carries ``{'episode': {'r', 'l'}, 'route_id'}`` when env i finishes an episode
    (tools/learn.py:121-126).  Observations are drawn on `device` (the rollout never visits the host); episodes end
    with probability ``1/mean_episode_len`` per step; the env reward is minus the squared steering command, only so
    that episode returns depend on the actions."""

    def __init__(self, num_envs: int, seed: int = 3, device="cpu", mean_episode_len: int = 400, routes=(0,)):
        self.num_envs, self.device = num_envs, torch.device(device)
        self.observation_space, self.metrics_space, self.action_space = _Space(OBS_SHAPE), _Space(METRICS_SHAPE), _Space(ACTION_SHAPE)
        self._g = _gen(seed, self.device.type if self.device.type == "cpu" else self.device)
        self._cpu = _gen(seed + 1)
        self._p_end = 1.0 / float(mean_episode_len)
        self._routes = tuple(routes)
        self._ret = [0.0] * num_envs
        self._len = [0] * num_envs
        self._route = [self._routes[i % len(self._routes)] for i in range(num_envs)]
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:     # tools/envs.py EnvEpoch.set_epoch
        self.epoch = epoch

    def _draw(self):
        return synth_obs(self.num_envs, self._g, self.device), synth_metrics(self.num_envs, self._g, self.device)

    def reset(self):
        self._ret = [0.0] * self.num_envs
        self._len = [0] * self.num_envs
        return self._draw()

    def step(self, action):
        a = torch.as_tensor(action).detach().float().reshape(self.num_envs, -1).cpu()
        obs, metrics = self._draw()
        rewards = -(a[:, :1] ** 2)
        done = (torch.rand(self.num_envs, generator=self._cpu) < self._p_end).tolist()
        infos = []
        for i in range(self.num_envs):
            self._ret[i] += float(rewards[i, 0])
            self._len[i] += 1
            info = {}
            if done[i]:
                info = {"episode": {"r": self._ret[i], "l": self._len[i]}, "route_id": self._route[i]}
                self._ret[i], self._len[i] = 0.0, 0
            infos.append(info)
        return obs, metrics, rewards, done, infos


class SyntheticEvalEnv:
    """Single evaluation env with a fixed episode length (tools/learn.py:225-252 protocol: unbatched obs / metrics,
    ``step(action[2]) -> (obs, metrics, reward, done, info)``, ``ep_length``)."""

    def __init__(self, ep_length: int = 16, seed: int = 5, device="cpu"):
        self.ep_length, self.device = ep_length, torch.device(device)
        self._g = _gen(seed, self.device.type if self.device.type == "cpu" else self.device)
        self._t, self._ret = 0, 0.0

    def reset(self):
        self._t, self._ret = 0, 0.0
        return synth_obs(1, self._g, self.device)[0], synth_metrics(1, self._g, self.device)[0]

    def step(self, action):
        self._t += 1
        self._ret += -float(action[0]) ** 2
        done = self._t >= self.ep_length - 1
        info = {"episode": {"r": self._ret, "l": self._t}} if done else {}
        return synth_obs(1, self._g, self.device)[0], synth_metrics(1, self._g, self.device)[0], 0.0, done, info
