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
from .graphs import StepGraph
from .optim import FusedClipAdam, world_size


class PPO():
    def __init__(self, actor_critic, clip_param, ppo_epoch, mini_batch_size, value_loss_coef, device, lr=None, eps=None,
                 betas=None, max_grad_norm=None, use_clipped_value_loss=True, gamma=None, decay=None, act_space=None):
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
        # Multi-GPU (torch.distributed initialised, envs sharded across ranks).  False: every rank
        # permutes its own env shard and `mini_batch_size` is the per-rank share of the global minibatch (equal shards,
        # mean of means).  True ("exact mode", needs rollouts.set_shard): one global permutation, global minibatches of
        # mini_batch_size * world rows, each rank processes the members it owns - bit-for-bit the sample sets, weights
        # and RNG stream of a single process holding all envs (tools/storage.py:60-66, algo/ppo.py:64-119).
        self.exact_sharding = False
        self._graph = StepGraph("PPO")
        self._dev_state = None

    def update(self, rollouts, expert_dataset=None):
        pol = self.actor_critic
        eng = pol.engine
        eng.sync_params()
        dev = eng.flat.flat.device
        T, N = rollouts.num_steps, rollouts.num_processes
        logstd = pol.base.logstd.tolist()
        act = pol.base.activation
        world = world_size()
        rank = dist.get_rank() if world > 1 else 0
        exact = bool(self.exact_sharding) and world > 1
        if exact and rollouts.shard != (rank, world):
            raise RuntimeError("exact_sharding needs rollouts.set_shard(rank, world) on every rank")
        self.optimizer.grad_scale = 1.0 if exact else None

        # advantage statistics over the whole rollout (algo/ppo.py:47-49); normalisation itself is fused into the loss
        stats = torch.empty(4, dtype=torch.float64, device=dev)      # {sum, sumsq, count} are written by the kernel call
        A.adv_stats(rollouts.returns, rollouts.value_preds, stats, T * N)
        if world > 1:
            dist.all_reduce(stats[:3], op=dist.ReduceOp.SUM)   # {sum, sumsq, count} are additive across env shards

        obs_rows, met_rows = rollouts.flat("obs"), rollouts.flat("metrics")
        act_rows, vp_rows = rollouts.flat("actions"), rollouts.flat("value_preds")
        ret_rows, lp_rows = rollouts.flat("returns"), rollouts.flat("action_log_probs")

        B = self.mini_batch_size
        Bg = B * world if exact else B            # rows one optimisation step averages over on this rank's scale
        use_bc = bool(expert_dataset)
        w_act = (1.0 - self.gamma) if use_bc else 1.0
        n_updates = 0
        n_bc_rows = 0
        # per-object device state the (replayable) step reads / writes: loss sums, advantage statistics, row indices
        # (kept per (batch size, device) for the object's lifetime: a captured graph holds their addresses)
        if self._dev_state is None:
            self._dev_state = {}
        if (B, str(dev)) not in self._dev_state:
            self._dev_state[(B, str(dev))] = (torch.zeros(4, dtype=torch.float64, device=dev), torch.zeros(4, dtype=torch.float64, device=dev),
                                              torch.zeros(B, dtype=torch.int64, device=dev))
        acc, stats_buf, idx_buf = self._dev_state[(B, str(dev))]
        A.zero_block(acc)
        if not eng.flat.grad_clean:        # a replayed step assumes the zeroed gradient buffer the previous step left behind
            eng.flat.grad.zero_(); eng.flat.grad_clean = True
        stats_buf.copy_(stats)
        opt = self.optimizer
        n_sched = self.ppo_epoch * ((T * N * (world if exact else 1)) // Bg)
        opt.begin_schedule(n_sched)
        clipped = bool(self.use_clipped_value_loss)
        clip, vcoef = float(self.clip_param), float(self.value_loss_coef)

        def device_step():
            """One optimisation step on the rows named by idx_buf: everything is enqueued on the current stream and reads
            only device-resident state, so it can be captured once and replayed (graphs.StepGraph)."""
            ws = eng.workspace(B)
            eng.load_inputs(obs_rows, met_rows, idx_buf, B)
            a_b = ws.buf("act", ws.rows, 2); vo_b = ws.buf("vold", ws.rows); r_b = ws.buf("ret", ws.rows)
            lp_b = ws.buf("olp", ws.rows)
            A.gather_rows(act_rows, idx_buf, a_b, B, 2, 2)
            A.gather_rows(vp_rows, idx_buf, vo_b, B, 1, 1)
            A.gather_rows(ret_rows, idx_buf, r_b, B, 1, 1)
            A.gather_rows(lp_rows, idx_buf, lp_b, B, 1, 1)
            head = eng.forward(B, training=True)
            d_head = ws.buf("dhead", ws.rows, 4)
            A.ppo_loss(head, a_b, lp_b, vo_b, r_b, None, stats_buf, d_head, None, None, acc, B, logstd, act, clip, vcoef, 1.0, 0,
                       clipped_value=clipped)
            eng.backward(B, d_head, reducer=opt.reducer)
            opt.step(from_device_hyper=True)
            eng.dirty = True
            eng.sync_params()

        def epoch_batches():   # one fresh permutation per epoch (tools/storage.py:60-63)
            if exact:
                return rollouts.sharded_minibatches(Bg)
            return ((None, idx) for idx in rollouts.minibatch_indices(B))

        for _ in range(self.ppo_epoch):
            for pos, idx in epoch_batches():
                opt.advance()
                if not use_bc and not exact:        # the common, fixed-shape step: replayed as one CUDA graph
                    idx_buf.copy_(idx, non_blocking=True)
                    ws = eng.workspace(B)
                    key = (ws.X0.data_ptr(), ws.rows, B, eng.flat.flat.data_ptr(), eng.flat.grad.data_ptr(), world, clipped, clip, vcoef,
                           tuple(logstd), act, self.max_grad_norm, opt.device_hyper(eng.flat).data_ptr()) + \
                        tuple(t.data_ptr() for t in (obs_rows, met_rows, act_rows, vp_rows, ret_rows, lp_rows, acc, stats_buf, idx_buf))
                    self._graph.run(key, device_step, dev)
                    n_updates += 1
                    continue
                Bl = int(idx.shape[0])             # rows of this step on this rank (== B unless exact sharding)
                Be = Be_norm = 0
                e_act = None
                if use_bc:   # algo/ppo.py:88-102: first batch of a fresh iterator over the expert loader
                    for e_batch in expert_dataset:
                        if exact:     # every rank sees the same expert batch: take an equal contiguous share of its rows
                            e_batch, Be_norm = _share_of(e_batch, rank, world)
                        e_obs, e_met, e_act_rows, e_idx, Be = expert_rows(e_batch, dev)
                        Be_norm = Be_norm or Be
                        break
                ws = eng.workspace(max(Bl + Be, 1))
                if Bl:
                    eng.load_inputs(obs_rows, met_rows, idx, Bl)
                if Be:
                    eng.load_inputs(e_obs, e_met, e_idx, Be, row0=Bl)
                    if e_idx is None:
                        e_act = e_act_rows
                    else:   # device-resident expert table: gather the batch's action rows
                        e_act = ws.buf("e_act", ws.rows, 2)
                        A.gather_rows(e_act_rows, e_idx, e_act, Be, 2, 2)
                if Bl + Be:
                    a_b = ws.buf("act", ws.rows, 2); vo_b = ws.buf("vold", ws.rows); r_b = ws.buf("ret", ws.rows)
                    lp_b = ws.buf("olp", ws.rows)
                    if Bl:
                        A.gather_rows(act_rows, idx, a_b, Bl, 2, 2)
                        A.gather_rows(vp_rows, idx, vo_b, Bl, 1, 1)
                        A.gather_rows(ret_rows, idx, r_b, Bl, 1, 1)
                        A.gather_rows(lp_rows, idx, lp_b, Bl, 1, 1)
                    head = eng.forward(Bl + Be, training=True)
                    d_head = ws.buf("dhead", ws.rows, 4)
                    if Bl:
                        A.ppo_loss(head, a_b, lp_b, vo_b, r_b, None, stats_buf, d_head, None, None, acc, Bl, logstd, act,
                                   clip, vcoef, float(w_act), 0, clipped_value=clipped, norm=Bg if exact else None)
                    if Be:
                        A.ppo_loss(head[Bl:], e_act, None, None, None, None, None, d_head[Bl:], None, None, acc, Be, logstd, act,
                                   0.0, 0.0, float(self.gamma), 1, norm=Be_norm if exact else None)
                    eng.backward(Bl + Be, d_head, reducer=opt.reducer)
                else:            # this rank owns no member of the global minibatch: it contributes a zero gradient
                    eng.zero_contribution(opt.reducer)
                n_bc_rows += Be_norm if exact else Be
                opt.step(from_device_hyper=True)
                eng.dirty = True
                eng.sync_params()
                n_updates += 1

        if world > 1:       # loss sums are additive across ranks (algo/ppo.py:128-141 reports global-batch means)
            dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        a = acc.cpu().tolist()                      # the only host synchronisation of the update
        rows = (Bg if exact else B * world) * n_updates
        value_loss = a[0] / rows
        gail_action_loss = a[1] / rows
        n_bc = n_bc_rows if exact else n_bc_rows * world
        bc_loss = (a[2] / n_bc) if n_bc else 0.0
        action_loss = self.gamma * bc_loss + (1 - self.gamma) * gail_action_loss if use_bc else gail_action_loss
        entropy = sum(0.5 + 0.5 * math.log(2 * math.pi) + v for v in logstd)
        if self.gamma is not None:
            self.gamma *= self.decay
        return value_loss, action_loss, entropy, bc_loss, gail_action_loss, self.gamma, logstd[0], logstd[1]


def _share_of(batch, rank: int, world: int):
    """Rank's contiguous share of an expert batch every rank holds a copy of -> (batch share, global row count)."""
    from .expert import DeviceBatch
    if isinstance(batch, DeviceBatch):
        n = batch.batch_rows
        lo, hi = n * rank // world, n * (rank + 1) // world
        return DeviceBatch(batch.obs_table, batch.metrics_table, batch.actions_table, batch.idx[lo:hi].contiguous()), n
    n = int(batch[0].shape[0])
    lo, hi = n * rank // world, n * (rank + 1) // world
    return tuple(t[lo:hi] for t in batch), n
