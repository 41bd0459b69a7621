"""The training iteration around the hot path: rollout collection with the policy on the
device, discriminator epochs with the reference's warm-up schedule, GAIL rewards, returns, PPO update, evaluation
episode, scalar log and checkpoint - the orchestration of tools/learn.py:89-306 with the storage, policy and critic
resident in HBM.

What is kept from the reference (so runs are comparable and checkpoints interchangeable):
  * environment protocol: ``envs.reset() -> (obs, metrics)``, ``envs.step(action) -> (obs, metrics, rewards, done,
    infos)``, ``observation_space / metrics_space / action_space`` with ``.shape`` (tools/envs.py vec-env API);
    ``infos[i]['episode'] = {'r', 'l'}`` and ``infos[i]['route_id']`` at episode ends (tools/learn.py:121-126);
  * ``run_params`` keys (params_variable.json): num_steps, num_env_steps, envs_params, routes, lr, use_linear_lr_decay,
    gail_epoch, gail_pre_epoch, gail_thre, gamma, gae_lambda, bcgail, eval_interval, log_interval, resume_training;
  * scalar names of tools/utli.py:9-88 (``ScalarLog``), the linear LR schedule (tools/utli.py:121-125), the GAIL-epoch
    warm-up (tools/learn.py:146-151), the checkpoint list ``[policy.state_dict(), disc.state_dict(), i_update, seconds]``
    (tools/learn.py:290-291, resume :81-87).
What changes: nothing moves between host and device inside an iteration except the simulator's observations
(``insert`` copies them into HBM as they arrive) and one read-back of rewards / masks for the episodic bookkeeping -
the reference's per-step ``.to(device)`` / ``.cpu()`` round trips and its ``discriminator.cpu()`` / ``actor_critic.cpu()``
shuffling have no counterpart because both networks fit in HBM next to the rollout.
"""
from __future__ import annotations

import math
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .storage import RolloutStorage

PPO_SCALARS = ("ppo_value", "ppo_loss", "ppo_entropy", "bc_loss", "gail_loss", "gail_gamma", "steer_std", "throttle_std")
DISC_SCALARS = ("dis_total_loss", "dis_policy_reward", "dis_expert_reward", "dis_loss", "dis_gp", "expert_loss", "policy_loss",
                "disc_pre_loss", "expert_pre_reward", "policy_pre_reward", "disc_after_loss", "expert_after_reward",
                "policy_after_reward")
TRAIN_SCALARS = ("Train reward", "Train steps", "Expert reward", "Eval steps", "Eval reward", "disc_eval_loss",
                 "expert_eval_reward", "policy_eval_reward")


class ScalarLog:
    """Scalars under the reference's tensorboard titles (tools/utli.py:9-100).  ``writer`` is anything with
    ``add_scalar(title, value, step)`` (a tensorboardX SummaryWriter in the reference); every record is also kept in
    ``history`` so a run can be inspected without tensorboard."""

    def __init__(self, writer=None):
        self.writer = writer
        self.history: List[Dict[str, float]] = []

    def record(self, titles: Sequence[str], values: Sequence, step: int) -> None:
        row = {"step": step}
        for t, v in zip(titles, values):
            v = float("nan") if v is None else float(v)
            row[t] = v
            if self.writer is not None:
                self.writer.add_scalar(t, v, step)
        self.history.append(row)

    def record_routes(self, routes_rewards: Dict[int, List[float]], step: int) -> None:
        titles, values = [], []
        for route, rs in routes_rewards.items():          # tools/utli.py:91-100
            if rs:
                titles += ["route_{:0>2d}_max_reward".format(route), "route_{:0>2d}_min_reward".format(route)]
                values += [max(rs), min(rs)]
        if titles:
            self.record(titles, values, step)


def linear_lr(initial_lr: float, update: int, total_updates: int) -> float:
    """tools/utli.py:121-125."""
    return initial_lr - initial_lr * (update / float(total_updates))


def set_lr(optimizer, lr: float) -> None:
    for group in optimizer.param_groups:
        group["lr"] = lr


def gail_epochs_for(update: int, gail_epoch: int, gail_pre_epoch: int, gail_thre: int) -> int:
    """Discriminator epochs of iteration `update` (1-based): linear warm-up from gail_pre_epoch (tools/learn.py:146-151)."""
    if update < gail_thre:
        return int(gail_epoch + (gail_pre_epoch - gail_epoch) * (gail_thre - (update - 1)) / gail_thre)
    return int(gail_epoch)


def episodic_gail_returns(gail_rewards: torch.Tensor, masks: torch.Tensor, carry: List[float]) -> List[float]:
    """Episode sums of the GAIL reward (tools/learn.py:204-209, a Python loop with one ``.item()`` per (step, env) in the
    reference): walking steps in order, ``masks[step][env] != 0`` adds the step's reward to the env's running sum,
    ``== 0`` closes the episode (its sum is reported, the reward of that step is dropped, the sum restarts at 0).
    ``carry`` holds the running sums across iterations and is updated in place.  One device read-back of two [T,N]
    tensors; the segment sums are formed per env with cumulative sums."""
    r = gail_rewards.detach().reshape(gail_rewards.shape[0], -1).double().cpu().numpy()
    m = masks.detach().reshape(masks.shape[0], -1)[: r.shape[0]].cpu().numpy() != 0
    T, N = r.shape
    closed = []                                   # (step, env, value) in the reference's append order
    for n in range(N):
        ends = np.flatnonzero(~m[:, n])
        contrib = np.where(m[:, n], r[:, n], 0.0)
        csum = np.concatenate([[0.0], np.cumsum(contrib)])
        start, run = 0, carry[n]
        for e in ends:
            closed.append((int(e), n, run + (csum[e] - csum[start])))
            start, run = e + 1, 0.0
        carry[n] = run + (csum[T] - csum[start])
    closed.sort(key=lambda t: (t[0], t[1]))
    return [v for _, _, v in closed]


def save_checkpoint(path, actor_critic, discriminator, i_update: int, seconds: float) -> None:
    torch.save([actor_critic.state_dict(), discriminator.state_dict(), i_update, seconds], path)


def load_checkpoint(path, actor_critic, discriminator):
    data = torch.load(path, map_location="cpu")
    actor_critic.load_state_dict(data[0])
    discriminator.load_state_dict(data[1])
    return int(data[2]), float(data[3])


def collect_rollout(envs, actor_critic, rollouts: RolloutStorage, episode_sink=None) -> None:
    """tools/learn.py:111-133 with the policy input read straight from the device-resident storage: one batched
    ``act`` for the N envs per step, ``insert`` copies the simulator's outputs into HBM."""
    dev = rollouts.obs.device
    for step in range(rollouts.num_steps):
        with torch.no_grad():
            value, action, action_log_prob = actor_critic.act(rollouts.obs[step], rollouts.metrics[step])
        obs, metrics, rewards, done, infos = envs.step(action)
        if episode_sink is not None:
            for info in infos:
                ep = info.get("episode") if isinstance(info, dict) else None
                if ep:
                    episode_sink(info)
        masks = torch.tensor([[0.0] if d else [1.0] for d in done], dtype=torch.float32)
        rollouts.insert(obs, metrics, action, action_log_prob, value, torch.as_tensor(rewards, dtype=torch.float32).reshape(-1, 1),
                        masks.to(dev, non_blocking=True))


def evaluation_episode(env_eval, actor_critic, rollout_eval: RolloutStorage):
    """Deterministic episode on the evaluation env (tools/learn.py:225-252); returns (steps, episode reward or None)."""
    dev = rollout_eval.obs.device
    obs, metrics = env_eval.reset()
    steps, reward, done = 0, None, False
    while not done and steps < rollout_eval.num_steps:
        o = torch.as_tensor(obs, dtype=torch.float32).to(dev).unsqueeze(0)
        m = torch.as_tensor(metrics, dtype=torch.float32).to(dev).unsqueeze(0)
        with torch.no_grad():
            _, actions, _ = actor_critic.act(o, m, deterministic=True)
        rollout_eval.obs[steps].copy_(o)
        rollout_eval.metrics[steps].copy_(m)
        rollout_eval.actions[steps].copy_(actions)
        obs, metrics, _, done, info = env_eval.step(actions.cpu().numpy()[0])
        steps += 1
        ep = info.get("episode") if isinstance(info, dict) else None
        if ep:
            reward = ep["r"]
    rollout_eval.obs[steps].copy_(torch.as_tensor(obs, dtype=torch.float32).to(dev).unsqueeze(0))
    rollout_eval.metrics[steps].copy_(torch.as_tensor(metrics, dtype=torch.float32).to(dev).unsqueeze(0))
    return steps, reward


def gail_learning(run_params: dict, envs, env_eval, actor_critic, agent, discriminator, gail_train_loader, gail_val_loader,
                  device, writer=None, model_path: str = "gail_model.pt", verbose: bool = False,
                  obs_dtype=torch.float32) -> ScalarLog:
    """The loop of tools/learn.py ``gailLearning_mujoco_origin``; returns the scalar log.
    ``obs_dtype=torch.uint8`` keeps the rollout observations in the byte store (storage.ByteObs; lossless for the simulator's
    uint8/255 observations, off-grid values raise)."""
    log = ScalarLog(writer)
    if torch.device(device).type == "cuda" and torch.device(device).index is not None:
        torch.cuda.set_device(torch.device(device))     # kernels and streams are issued on the current device
    nenv = len(run_params["envs_params"])
    nbatch = int(math.floor(run_params["num_steps"] / nenv))
    nupdates = int(math.floor(run_params["num_env_steps"] / run_params["num_steps"]))
    rollouts = RolloutStorage(nbatch, nenv, envs.observation_space.shape, envs.metrics_space.shape, envs.action_space.shape,
                              device=device, obs_dtype=obs_dtype)
    rollout_eval = None
    if env_eval is not None:
        rollout_eval = RolloutStorage(env_eval.ep_length, 1, envs.observation_space.shape, envs.metrics_space.shape,
                                      envs.action_space.shape, device=device)
    actor_critic.to(device)
    discriminator.to(device)
    carry = [0.0] * nenv
    obs, metrics = envs.reset()
    rollouts.obs[0].copy_(obs)
    rollouts.metrics[0].copy_(metrics)
    start = time.time()
    i_update, time_step = 0, 0
    steps_eval, eval_reward = 0, None
    disc_eval = (float("nan"),) * 3
    if run_params.get("resume_training") and os.path.exists(model_path):
        i_update, elapsed = load_checkpoint(model_path, actor_critic, discriminator)
        start -= elapsed

    while i_update < nupdates:
        i_update += 1
        episodes: List[dict] = []
        routes_rewards: Dict[int, List[float]] = {r: [] for r in run_params.get("routes", [])}

        def sink(info):
            episodes.append(info["episode"])
            routes_rewards.setdefault(info.get("route_id", 0), []).append(info["episode"]["r"])

        if run_params.get("use_linear_lr_decay"):
            set_lr(agent.optimizer, linear_lr(run_params["lr"], i_update, nupdates))
        if hasattr(envs, "set_epoch"):
            envs.set_epoch(i_update)
        collect_rollout(envs, actor_critic, rollouts, sink)
        time_step += nbatch
        with torch.no_grad():
            rollouts.value_preds[-1] = actor_critic.get_value(rollouts.obs[-1], rollouts.metrics[-1])

        pre = discriminator.compute_loss(gail_val_loader, rollouts)
        n_epochs = gail_epochs_for(i_update, run_params["gail_epoch"], run_params.get("gail_pre_epoch", run_params["gail_epoch"]),
                                   run_params.get("gail_thre", 0))
        d_out = [discriminator.update(gail_train_loader, rollouts) for _ in range(n_epochs)]
        post = discriminator.compute_loss(gail_val_loader, rollouts)
        d_mean = [float(np.mean([float(t[j]) for t in d_out])) if d_out else float("nan") for j in range(7)]
        log.record(DISC_SCALARS, d_mean + [float(v) for v in pre] + [float(v) for v in post], i_update)

        discriminator.predict_rewards_rollout(rollouts)          # tools/learn.py:196-202, one batched pass
        epgail = episodic_gail_returns(rollouts.gail_rewards, rollouts.masks, carry)
        rollouts.compute_returns(run_params["gamma"], run_params["gae_lambda"])
        p_out = agent.update(rollouts, gail_train_loader if run_params.get("bcgail") else None)

        if env_eval is not None and (i_update % run_params.get("eval_interval", 1) == 0 or eval_reward is None):
            steps_eval, eval_reward = evaluation_episode(env_eval, actor_critic, rollout_eval)
            if steps_eval > 1:
                disc_eval = discriminator.compute_loss(gail_val_loader, rollout_eval, batch_size=steps_eval - 1)
        log.record(PPO_SCALARS, p_out, i_update)
        rollouts.after_update()
        if not episodes:
            continue
        eprew = float(np.mean([e["r"] for e in episodes]))
        eplen = float(np.mean([e["l"] for e in episodes]))
        log.record(TRAIN_SCALARS, (eprew, eplen, float(np.mean(epgail)) if epgail else float("nan"), steps_eval, eval_reward,
                                   *disc_eval), i_update)
        log.record_routes(routes_rewards, i_update)
        save_checkpoint(model_path, actor_critic, discriminator, i_update, time.time() - start)
        if verbose:
            print("Episode: %d,   Time steps: %d,   Mean length: %d    Mean Reward: %f    Mean Gail Reward:%f"
                  % (i_update, time_step, eplen, eprew, float(np.mean(epgail)) if epgail else float("nan")))
    return log
