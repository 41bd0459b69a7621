"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional, torch-CPU fp32 restatement of the gail-carla learning hot path.
Every function names the reference lines it follows (paths relative to
/root/reference).  State lives in plain dicts keyed by the reference's
``state_dict`` names, so outputs can be compared tensor by tensor with the
reference modules (tests/golden/make_golden.py does that) and with the CUDA
path (tests/test_*_gpu.py).

Pinned: yes - against the unmodified reference executed in the build container;
see tests/golden/make_golden.py and tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

OBS_SHAPE = (3, 192, 192)
NORM_MEAN = (0.485, 0.456, 0.406)   # tools/model.py:154
NORM_STD = (0.229, 0.224, 0.225)    # tools/model.py:155
LRELU = 0.2                         # tools/model.py:138-144
CONV_CH = (3, 32, 64, 128, 256)     # tools/model.py:137-143
FEAT = 25600                        # tools/model.py:148-153  (256*10*10)
N_METRIC_FEAT = 13                  # tools/model.py:171-177


# --------------------------------------------------------------------------- #
# parameter construction: same module construction order as the reference so   #
# the default-generator draws line up (tools/model.py:56-69,131-155,167-177;    #
# algo/wdgail.py:19-33)                                                         #
# --------------------------------------------------------------------------- #
def _conv_stack(prefix: str, out: Params) -> None:
    for i in range(4):
        m = nn.Conv2d(CONV_CH[i], CONV_CH[i + 1], 4, stride=2)
        out[f"{prefix}main.{2 * i}.weight"] = m.weight.detach().clone()
        out[f"{prefix}main.{2 * i}.bias"] = m.bias.detach().clone()


def _linear(name: str, fin: int, fout: int, out: Params) -> None:
    m = nn.Linear(fin, fout)
    out[name + ".weight"] = m.weight.detach().clone()
    out[name + ".bias"] = m.bias.detach().clone()


def init_policy_params() -> Params:
    """tools/model.py:15-23,56-69 - draws from torch's default CPU generator."""
    p: Params = {}
    _conv_stack("base.obs_processor.", p)
    p["base.metrics_processor.road_option_embedding.weight"] = nn.Embedding(10, 8).weight.detach().clone()
    _linear("base.body.body.0", FEAT + N_METRIC_FEAT, 512, p)
    _linear("base.body.body.2", 512, 512, p)
    _linear("base.body.body.4", 512, 512, p)
    _linear("base.head.head.0", 512, 256, p)
    _linear("base.head.head.2", 256, 3, p)
    return p


def init_disc_params(hidden_dim: int = 100) -> Params:
    """algo/wdgail.py:19-33."""
    p: Params = {}
    _conv_stack("obs_processor.", p)
    p["metrics_processor.road_option_embedding.weight"] = nn.Embedding(10, 8).weight.detach().clone()
    _linear("trunk.0", FEAT + N_METRIC_FEAT + 2, hidden_dim, p)
    _linear("trunk.2", hidden_dim, 1, p)
    return p


# --------------------------------------------------------------------------- #
# trunks                                                                        #
# --------------------------------------------------------------------------- #
def obs_features(p: Params, prefix: str, obs: torch.Tensor) -> torch.Tensor:
    """tools/model.py:157-164: per-channel normalise, 4x conv(k4,s2)+LeakyReLU, flatten (NCHW order)."""
    mean = torch.tensor(NORM_MEAN, dtype=obs.dtype, device=obs.device).view(1, 3, 1, 1)
    std = torch.tensor(NORM_STD, dtype=obs.dtype, device=obs.device).view(1, 3, 1, 1)
    x = (obs - mean) / std
    for i in range(4):
        x = F.conv2d(x, p[f"{prefix}main.{2 * i}.weight"], p[f"{prefix}main.{2 * i}.bias"], stride=2)
        x = F.leaky_relu(x, LRELU)
    return x.reshape(x.shape[0], -1)


def metrics_features(emb: torch.Tensor, metrics: torch.Tensor) -> torch.Tensor:
    """tools/model.py:179-213: [x,y,v,c] -> [1000x,1000y,1000r,0.3*atan2(y,x),0.1v] ++ Embedding[int(c)].

    The reference does the feature arithmetic in numpy fp32 on the host; so do we.
    """
    m = metrics.detach().cpu().numpy()
    x, y = m[:, 0], m[:, 1]
    r = np.sqrt(x * x + y * y)
    th = np.arctan2(y, x)
    cols = [1000 * torch.from_numpy(x).float(), 1000 * torch.from_numpy(y).float(),
            1000 * torch.from_numpy(r).float(), 0.3 * torch.from_numpy(th).float(),
            0.1 * torch.from_numpy(m[:, 2]).float()]
    head = torch.stack(cols, dim=1).to(metrics.device)          # tools/model.py:207 .to(metrics.device)
    idx = torch.from_numpy(m[:, 3]).long().to(metrics.device)   # tools/model.py:204
    return torch.cat([head, emb[idx]], dim=1)


def policy_base(p: Params, obs, metrics, activation: bool, logstd: Sequence[float]):
    """tools/model.py:71-86 (CNNBase.forward) + :89-128 (NNBody/NNHead)."""
    f = torch.cat([obs_features(p, "base.obs_processor.", obs),
                   metrics_features(p["base.metrics_processor.road_option_embedding.weight"], metrics)], dim=1)
    h = f
    for k in ("base.body.body.0", "base.body.body.2", "base.body.body.4"):
        h = F.leaky_relu(F.linear(h, p[k + ".weight"], p[k + ".bias"]), LRELU)
    h = F.leaky_relu(F.linear(h, p["base.head.head.0.weight"], p["base.head.head.0.bias"]), LRELU)
    out = F.linear(h, p["base.head.head.2.weight"], p["base.head.head.2.bias"])
    value = out[:, 0:1]
    mu = out[:, 1:]
    if activation:  # tools/model.py:80-82
        mu = torch.stack([torch.tanh(mu[:, 0]), torch.sigmoid(mu[:, 1])], dim=1)
    ls = torch.tensor(list(logstd), dtype=mu.dtype, device=mu.device).view(1, 2).expand_as(mu)
    return value, mu, ls


def normal_log_prob(mu, logstd, action):
    """torch.distributions.Normal.log_prob summed over the action dims (tools/model.py:47-49)."""
    var = torch.exp(logstd) ** 2
    lp = -((action - mu) ** 2) / (2 * var) - logstd - math.log(math.sqrt(2 * math.pi))
    return lp.sum(-1, keepdim=True)


def normal_entropy(logstd):
    """Normal.entropy().sum(-1).mean() (tools/model.py:50)."""
    return (0.5 + 0.5 * math.log(2 * math.pi) + logstd).sum(-1).mean()


def evaluate_actions(p: Params, obs, metrics, action, activation=True, logstd=(-1.4, -3.2)):
    """tools/model.py:45-53."""
    value, mu, ls = policy_base(p, obs, metrics, activation, logstd)
    return value, normal_log_prob(mu, ls, action), normal_entropy(ls), ls[0, 0].detach(), ls[0, 1].detach()


def act_deterministic(p: Params, obs, metrics, activation=True, logstd=(-1.4, -3.2)):
    """tools/model.py:25-36 with deterministic=True."""
    value, mu, ls = policy_base(p, obs, metrics, activation, logstd)
    return value, mu, normal_log_prob(mu, ls, mu)


def disc_forward(p: Params, obs, metrics, action) -> torch.Tensor:
    """algo/wdgail.py:40-54."""
    f = torch.cat([obs_features(p, "obs_processor.", obs),
                   metrics_features(p["metrics_processor.road_option_embedding.weight"], metrics), action], dim=1)
    h = F.leaky_relu(F.linear(f, p["trunk.0.weight"], p["trunk.0.bias"]), LRELU)
    return F.linear(h, p["trunk.2.weight"], p["trunk.2.bias"])


def predict_reward(p: Params, obs, metrics, action) -> torch.Tensor:
    """algo/wdgail.py:181-189: -log(1 - sigmoid(D)); gamma/masks/update_rms are ignored by the reference."""
    with torch.no_grad():
        d = disc_forward(p, obs, metrics, action)
        return -(1 - torch.sigmoid(d)).log()


# --------------------------------------------------------------------------- #
# rollout maths                                                                 #
# --------------------------------------------------------------------------- #
def gae_returns(gail_rewards, value_preds, masks, gamma: float, gae_lambda: float) -> torch.Tensor:
    """tools/storage.py:37-50.  Inputs [T,N,1] / [T+1,N,1]; returns [T+1,N,1] with row T left at 0."""
    T = gail_rewards.shape[0]
    ret = torch.zeros_like(value_preds)
    gae = torch.zeros_like(value_preds[0])
    for t in reversed(range(T)):
        delta = gail_rewards[t] + gamma * value_preds[t + 1] * masks[t + 1] - value_preds[t]
        gae = delta + gamma * gae_lambda * masks[t + 1] * gae
        ret[t] = gae + value_preds[t]
    return ret


def normalized_advantages(returns, value_preds) -> torch.Tensor:
    """algo/ppo.py:47-49 (unbiased std, +1e-5)."""
    adv = returns[:-1] - value_preds[:-1]
    return (adv - adv.mean()) / (adv.std() + 1e-5)


def minibatch_indices(batch_size: int, mini_batch_size: int) -> Iterator[List[int]]:
    """tools/storage.py:60-63: BatchSampler(SubsetRandomSampler(range(n)), mb, drop_last=True).

    SubsetRandomSampler.__iter__ draws torch.randperm(n) from the default CPU generator.
    """
    perm = torch.randperm(batch_size).tolist()
    for s in range(0, batch_size - mini_batch_size + 1, mini_batch_size):
        yield perm[s:s + mini_batch_size]


def ppo_losses(values, logp, old_logp, adv, value_old, returns, clip: float):
    """algo/ppo.py:80-85,104-111."""
    ratio = torch.exp(logp - old_logp)
    action_loss = -torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv).mean()
    vclip = value_old + (values - value_old).clamp(-clip, clip)
    value_loss = 0.5 * torch.max((values - returns) ** 2, (vclip - returns) ** 2).mean()
    return value_loss, action_loss


# --------------------------------------------------------------------------- #
# optimiser pieces (torch.nn.utils.clip_grad_norm_ + torch.optim.Adam restated) #
# --------------------------------------------------------------------------- #
def clip_grad_norm(grads: Iterable[torch.Tensor], max_norm: float) -> float:
    """clip_grad_norm_ as called at algo/ppo.py:116-117, algo/wdgail.py:142-143."""
    grads = list(grads)
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return float(total)


class AdamState:
    """torch.optim.Adam (no amsgrad / weight decay), as built at algo/ppo.py:43, algo/wdgail.py:35."""

    def __init__(self, params: Params, lr: float, eps: float, betas: Tuple[float, float]):
        self.lr, self.eps, self.b1, self.b2 = lr, eps, betas[0], betas[1]
        self.t = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def step(self, params: Params, grads: Params) -> None:
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        for k, p in params.items():
            g = grads[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            p.data.addcdiv_(self.m[k], denom, value=-self.lr / bc1)


def _leaf(params: Params) -> Params:
    return {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}


# --------------------------------------------------------------------------- #
# PPO.update                                                                    #
# --------------------------------------------------------------------------- #
def ppo_update(params: Params, adam: AdamState, ro: dict, *, clip_param: float, ppo_epoch: int,
               mini_batch_size: int, value_loss_coef: float, max_grad_norm: float,
               activation=True, logstd=(-1.4, -3.2), expert_loader=None, bc_gamma=None, decay=None,
               use_clipped_value_loss=True):
    """algo/ppo.py:45-141.  ``ro`` holds the RolloutStorage tensors by attribute name.

    Updates ``params`` in place; returns (8-tuple like the reference, new bc_gamma).
    """
    T, N = ro["gail_rewards"].shape[:2]
    adv = normalized_advantages(ro["returns"], ro["value_preds"]).view(-1, 1)
    obs = ro["obs"][:-1].reshape(T * N, *ro["obs"].shape[2:])
    met = ro["metrics"][:-1].reshape(T * N, -1)
    act = ro["actions"].reshape(T * N, -1)
    vold = ro["value_preds"][:-1].reshape(-1, 1)
    ret = ro["returns"][:-1].reshape(-1, 1)
    olp = ro["action_log_probs"].reshape(-1, 1)
    acc = dict(v=0.0, a=0.0, ent=0.0, bc=0.0, ga=0.0, s=0.0, th=0.0)
    n_updates = 0
    for _ in range(ppo_epoch):
        for idx in minibatch_indices(T * N, mini_batch_size):
            leaf = _leaf(params)
            values, logp, ent, s_std, t_std = evaluate_actions(leaf, obs[idx], met[idx], act[idx], activation, logstd)
            value_loss, action_loss = ppo_losses(values, logp, olp[idx], adv[idx], vold[idx], ret[idx], clip_param)
            if not use_clipped_value_loss:  # algo/ppo.py:112-113
                value_loss = 0.5 * (ret[idx] - values).pow(2).mean()
            acc["ga"] += action_loss.item()
            if expert_loader:  # algo/ppo.py:88-102: first batch of a fresh iterator
                for e_obs, e_met, e_act in expert_loader:
                    _, e_logp, _, _, _ = evaluate_actions(leaf, e_obs, e_met, e_act, activation, logstd)
                    bcloss = -e_logp.mean()
                    acc["bc"] += bcloss.item()
                    action_loss = bc_gamma * bcloss + (1 - bc_gamma) * action_loss
                    break
            (value_loss * value_loss_coef + action_loss).backward()
            grads = {k: v.grad for k, v in leaf.items()}
            clip_grad_norm(grads.values(), max_grad_norm)
            adam.step(params, grads)
            acc["v"] += value_loss.item(); acc["a"] += action_loss.item(); acc["ent"] += ent.item()
            acc["s"] += s_std.item(); acc["th"] += t_std.item()
            n_updates += 1
    for k in acc:
        acc[k] /= n_updates
    if bc_gamma is not None:
        bc_gamma *= decay
    return (acc["v"], acc["a"], acc["ent"], acc["bc"], acc["ga"], bc_gamma, acc["s"], acc["th"])


# --------------------------------------------------------------------------- #
# Discriminator.update / compute_loss                                           #
# --------------------------------------------------------------------------- #
def grad_penalty(leaf: Params, e, pbatch, alpha: torch.Tensor, lambda_: float = 10.0):
    """algo/wdgail.py:56-98 given the alpha draw ``torch.rand(B,1,1,1)`` (:66).

    Only the gradient w.r.t. the (pre-normalisation) image input is used (:85-91 ``[0]``).
    Note metrics are mixed *before* ProcessMetrics, so the road option is int(alpha*c_e+(1-alpha)*c_p).
    """
    B = e[0].shape[0]
    alpha = alpha.to(e[0].device)           # algo/wdgail.py:68 (.to(expert_state.device))
    a2 = alpha.view(B, 1)
    x = (alpha * e[0] + (1 - alpha) * pbatch[0]).detach().requires_grad_(True)
    m = a2 * e[1] + (1 - a2) * pbatch[1]
    a = a2 * e[2] + (1 - a2) * pbatch[2]
    d = disc_forward(leaf, x, m, a)
    g = torch.autograd.grad(d, x, torch.ones_like(d), create_graph=True, retain_graph=True)[0]
    g = g.view(B, -1)
    return lambda_ * ((g.norm(2, dim=1) - 1) ** 2).mean()


def policy_batches(ro: dict, mini_batch_size: int, batch_size: Optional[int] = None):
    """tools/storage.py:52-79 restricted to what the discriminator reads (obs, metrics, actions)."""
    T, N = ro["gail_rewards"].shape[:2]
    obs = ro["obs"][:-1].reshape(T * N, *ro["obs"].shape[2:])
    met = ro["metrics"][:-1].reshape(T * N, -1)
    act = ro["actions"].reshape(T * N, -1)
    for idx in minibatch_indices(batch_size or T * N, mini_batch_size):
        yield obs[idx], met[idx], act[idx]


def disc_update(params: Params, adam: AdamState, expert_loader, ro: dict, max_grad_norm: float):
    """algo/wdgail.py:100-147.  Returns the reference's 7-tuple."""
    tot = dict(loss=0.0, pr=0.0, er=0.0, g=0.0, gp=0.0, ea=0.0, pa=0.0)
    n = 0
    # zip(expert_loader, generator): the expert iterator is advanced first (algo/wdgail.py:112)
    for e, pb in zip(expert_loader, policy_batches(ro, expert_loader.batch_size)):
        leaf = _leaf(params)
        pd = disc_forward(leaf, *pb)
        ed = disc_forward(leaf, *e)
        tot["pr"] += pd.sum().item(); tot["er"] += ed.sum().item()
        el, pl = torch.tanh(ed).mean(), torch.tanh(pd).mean()
        B = pb[0].shape[0]
        wd = el - pl
        alpha = torch.rand(B, 1, 1, 1)
        gp = grad_penalty(leaf, e, pb, alpha)
        tot["ea"] += el.item() * B; tot["pa"] += pl.item() * B
        tot["loss"] += (-wd + gp).item() * B; tot["g"] += wd.item() * B; tot["gp"] += gp.item() * B
        n += B
        (-wd + gp).backward()
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
        clip_grad_norm(grads.values(), max_grad_norm)
        adam.step(params, grads)
    return (tot["loss"] / n, tot["pr"] / n, tot["er"] / n, tot["g"] / n, tot["gp"] / n, tot["ea"] / n, tot["pa"] / n)


def disc_compute_loss(params: Params, expert_loader, ro: dict, batch_size: Optional[int] = None):
    """algo/wdgail.py:149-179."""
    total = er = pr = 0.0
    n = 0
    with torch.no_grad():
        for e, pb in zip(expert_loader, policy_batches(ro, expert_loader.batch_size, batch_size)):
            el, pl = torch.tanh(disc_forward(params, *e)), torch.tanh(disc_forward(params, *pb))
            er += el.sum().item(); pr += pl.sum().item(); total += (el - pl).sum().item()
            n += pb[0].shape[0]
    if n == 0:
        return total, er, pr
    return total / n, er / n, pr / n


# --------------------------------------------------------------------------- #
# RunningMeanStd                                                                #
# --------------------------------------------------------------------------- #
def rms_merge(mean, var, count, batch_mean, batch_var, batch_count):
    """common/running_mean_std.py:20-31 (Chan et al. parallel merge, float64)."""
    delta = batch_mean - mean
    tot = count + batch_count
    new_mean = mean + delta * batch_count / tot
    m2 = var * count + batch_var * batch_count + np.square(delta) * count * batch_count / tot
    return new_mean, m2 / tot, tot


def rms_update(state, x: np.ndarray):
    """common/running_mean_std.py:10-14: population variance over axis 0."""
    mean, var, count = state
    return rms_merge(mean, var, count, np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])
