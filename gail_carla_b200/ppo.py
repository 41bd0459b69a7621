"""Drop-in for algo/ppo.py ``PPO``: same constructor, ``update(rollouts, expert_dataset=None)`` returning the same
8-tuple, ``.optimizer`` with mutable ``param_groups``.

One optimisation step (algo/ppo.py:64-119) is: index-gather of the minibatch straight out of the HBM-resident
storage -> trunk forward (tcgen05) -> fused head-tail + PPO-loss forward/backward kernel -> hand-derived trunk
backward (tcgen05 dgrad/wgrad) -> [NCCL all-reduce] -> fused clip-grad-norm + Adam.  No autograd, no per-minibatch
host synchronisation: loss terms accumulate in device doubles and are read back once per ``update``.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import _abi as A
from .expert import expert_rows
from .optim import FusedClipAdam


class PPO():
    def __init__(self, actor_critic, clip_param, ppo_epoch, mini_batch_size, value_loss_coef, device, lr=None, eps=None,
                 betas=None, max_grad_norm=None, use_clipped_value_loss=True, gamma=None, decay=None, act_space=None):
        if not use_clipped_value_loss:
            raise NotImplementedError("the reference always trains with use_clipped_value_loss=True (wdail_carla.py:211-224)")
        self.actor_critic = actor_critic
        self.clip_param = clip_param
        self.ppo_epoch = ppo_epoch
        self.mini_batch_size = mini_batch_size
        self.act_space = act_space
        self.value_loss_coef = value_loss_coef
        self.device = device
        self.gamma = gamma          # BC mixing weight (gailgamma), algo/ppo.py:37,99
        self.decay = decay
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.optimizer = FusedClipAdam(lambda: actor_critic.engine.flat, actor_critic.parameters(), lr, eps, betas, max_grad_norm)

    def update(self, rollouts, expert_dataset=None):
        pol = self.actor_critic
        eng = pol.engine
        eng.sync_params()
        dev = eng.flat.flat.device
        T, N = rollouts.num_steps, rollouts.num_processes
        logstd = pol.base.logstd.tolist()
        act = pol.base.activation

        # advantage statistics over the whole rollout (algo/ppo.py:47-49); normalisation itself is fused into the loss
        stats = torch.zeros(4, dtype=torch.float64, device=dev)
        A.adv_stats(rollouts.returns, rollouts.value_preds, stats, T * N)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(stats[:3], op=dist.ReduceOp.SUM)   # {sum, sumsq, count} are additive across env shards

        obs_rows, met_rows = rollouts.flat("obs"), rollouts.flat("metrics")
        act_rows, vp_rows = rollouts.flat("actions"), rollouts.flat("value_preds")
        ret_rows, lp_rows = rollouts.flat("returns"), rollouts.flat("action_log_probs")

        B = self.mini_batch_size
        use_bc = bool(expert_dataset)
        w_act = (1.0 - self.gamma) if use_bc else 1.0
        acc = torch.zeros(4, dtype=torch.float64, device=dev)
        n_updates = 0
        n_bc_rows = 0
        for _ in range(self.ppo_epoch):
            for idx in rollouts.minibatch_indices(B):
                Be = 0
                e_act = None
                if use_bc:   # algo/ppo.py:88-102: first batch of a fresh iterator over the expert loader
                    for e_batch in expert_dataset:
                        e_obs, e_met, e_act_rows, e_idx, Be = expert_rows(e_batch, dev)
                        break
                ws = eng.workspace(B + Be)
                eng.load_inputs(obs_rows, met_rows, idx, B)
                if Be:
                    eng.load_inputs(e_obs, e_met, e_idx, Be, row0=B)
                    if e_idx is None:
                        e_act = e_act_rows
                    else:   # device-resident expert table: gather the batch's action rows
                        e_act = ws.buf("e_act", ws.rows, 2)
                        A.gather_rows(e_act_rows, e_idx, e_act, Be, 2, 2)
                a_b = ws.buf("act", ws.rows, 2); vo_b = ws.buf("vold", ws.rows); r_b = ws.buf("ret", ws.rows)
                lp_b = ws.buf("olp", ws.rows)
                A.gather_rows(act_rows, idx, a_b, B, 2, 2)
                A.gather_rows(vp_rows, idx, vo_b, B, 1, 1)
                A.gather_rows(ret_rows, idx, r_b, B, 1, 1)
                A.gather_rows(lp_rows, idx, lp_b, B, 1, 1)
                head = eng.forward(B + Be, training=True)
                d_head = ws.buf("dhead", ws.rows, 4)
                A.ppo_loss(head, a_b, lp_b, vo_b, r_b, None, stats, d_head, None, None, acc, B, logstd, act,
                           float(self.clip_param), float(self.value_loss_coef), float(w_act), 0)
                if Be:
                    A.ppo_loss(head[B:], e_act, None, None, None, None, None, d_head[B:], None, None, acc, Be, logstd, act,
                               0.0, 0.0, float(self.gamma), 1)
                    n_bc_rows += Be
                eng.backward(B + Be, d_head)
                self.optimizer.step()
                eng.dirty = True
                eng.sync_params()
                n_updates += 1

        a = acc.cpu().tolist()                      # the only host synchronisation of the update
        value_loss = a[0] / (B * n_updates)
        gail_action_loss = a[1] / (B * n_updates)
        bc_loss = (a[2] / n_bc_rows) if n_bc_rows else 0.0
        action_loss = self.gamma * bc_loss + (1 - self.gamma) * gail_action_loss if use_bc else gail_action_loss
        entropy = sum(0.5 + 0.5 * math.log(2 * math.pi) + v for v in logstd)
        if self.gamma is not None:
            self.gamma *= self.decay
        return value_loss, action_loss, entropy, bc_loss, gail_action_loss, self.gamma, logstd[0], logstd[1]
