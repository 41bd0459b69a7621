"""Behaviour-cloning pre-training: learn_bc.py:15-72 of the reference on the B200 trunk.

Reference loop: Adam(lr=3e-4, default eps / betas, *no* gradient clipping - ``max_grad_norm`` is defined but never used),
per expert batch ``loss = -evaluate_actions(...).log_probs.mean()`` (``ent_weight = 0``), epoch loss = mean of the batch
losses, evaluation loss over ``eval_loader`` without gradients, scalars ``loss`` / ``eval_loss`` per epoch, checkpoint of
``actor_critic.state_dict()`` whenever the evaluation loss improves.  Here each batch is one forward pass of the policy
trunk, one launch of the fused head kernel in BC mode (``gc_ppo_loss_fwd_bwd`` mode 1: -mean log-prob and its gradient
w.r.t. the head outputs), the hand-derived trunk backward and the fused Adam step; batches may be the reference's host
tensors or ``DeviceBatch`` es of a device-resident expert table (no PCIe traffic at all in that case).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import torch

from . import _abi as A
from .expert import expert_rows
from .optim import FusedClipAdam


def _bc_batch(actor_critic, batch, train: bool, optimizer: Optional[FusedClipAdam]) -> Tuple[torch.Tensor, int]:
    """One expert batch -> (device scalar holding sum of -log pi(a|s) over the batch, rows)."""
    eng = actor_critic.engine
    eng.sync_params()
    dev = eng.flat.flat.device
    obs, met, act_rows, idx, B = expert_rows(batch, dev)
    ws = eng.workspace(B)
    eng.load_inputs(obs, met, idx, B)
    if idx is None:
        actions = act_rows
    else:
        actions = ws.buf("e_act", ws.rows, 2)
        A.gather_rows(act_rows, idx, actions, B, 2, 2)
    head = eng.forward(B, training=train)
    acc = torch.zeros(4, dtype=torch.float64, device=dev)
    logstd = actor_critic.base.logstd.tolist()
    if train:
        d_head = ws.buf("dhead", ws.rows, 4)
        A.ppo_loss(head, actions, None, None, None, None, None, d_head, None, None, acc, B, logstd, actor_critic.base.activation,
                   0.0, 0.0, 1.0, 1)
        eng.backward(B, d_head)
        optimizer.step()
        eng.dirty = True
        eng.sync_params()
        return acc[2], B
    logp = ws.buf("bc_logp", ws.rows)
    value = ws.buf("bc_value", ws.rows)
    A.ppo_loss(head, actions, None, None, None, None, None, None, value, logp, None, B, logstd, actor_critic.base.activation,
               0.0, 0.0, 0.0, 2)
    return -logp[:B].double().sum(), B


def learn_bc(actor_critic, device, expert_loader, eval_loader, episodes: int = 300, lr: float = 3e-4, writer=None,
             save_path: Optional[str] = "carla_actor_bc.pt") -> List[Tuple[float, float]]:
    """learn_bc.py:15-72; returns [(loss, eval_loss)] per epoch."""
    actor_critic.to(device)
    optimizer = FusedClipAdam(lambda: actor_critic.engine.flat, actor_critic.parameters(), lr, 1e-8, (0.9, 0.999), None)
    history, best = [], math.inf
    with torch.no_grad():
        for epoch in range(episodes):
            train_terms, eval_terms = [], []
            for batch in expert_loader:
                s, n = _bc_batch(actor_critic, batch, True, optimizer)
                train_terms.append(s / n)                    # batch loss = -mean log-prob (kept on the device)
            for batch in eval_loader:
                s, n = _bc_batch(actor_critic, batch, False, None)
                eval_terms.append(s / n)
            # one read-back per epoch
            loss = float(torch.stack(train_terms).mean().item()) if train_terms else float("nan")
            eval_loss = float(torch.stack(eval_terms).mean().item()) if eval_terms else float("nan")
            if writer is not None:
                writer.add_scalar("loss", loss, epoch)
                writer.add_scalar("eval_loss", eval_loss, epoch)
            history.append((loss, eval_loss))
            if best > eval_loss:
                if save_path:
                    torch.save(actor_critic.state_dict(), save_path)
                best = eval_loss
    return history
