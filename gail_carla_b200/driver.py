"""The env-free slice of the reference's training iteration (tools/learn.py:137-223,269): everything that happens
between the end of rollout collection and the next rollout.  Used by bench.py, the parity tests and smoke()."""
from __future__ import annotations

import torch


def update_iteration(actor_critic, agent, discriminator, rollouts, expert_loader, *, gamma, gae_lambda, gail_epoch=1,
                     bcgail=False, diagnostics=False, batched_rewards=True):
    """Returns (disc_tuples, ppo_tuple[, loss_before, loss_after])."""
    # tools/learn.py:137-139  bootstrap value
    rollouts.value_preds[-1] = actor_critic.get_value(rollouts.obs[-1], rollouts.metrics[-1])
    extra = []
    if diagnostics:   # tools/learn.py:144-145
        extra.append(discriminator.compute_loss(expert_loader, rollouts))
    # tools/learn.py:159-169  discriminator epochs
    d_out = [discriminator.update(expert_loader, rollouts) for _ in range(gail_epoch)]
    if diagnostics:   # tools/learn.py:178-179
        extra.append(discriminator.compute_loss(expert_loader, rollouts))
    # tools/learn.py:196-202  GAIL rewards for every stored step
    if batched_rewards:
        discriminator.predict_rewards_rollout(rollouts)
    else:
        for step in range(rollouts.num_steps):
            rollouts.gail_rewards[step] = discriminator.predict_reward(
                rollouts.obs[step], rollouts.metrics[step], rollouts.actions[step], gamma, rollouts.masks[step]
            ).to(rollouts.gail_rewards.device)
    rollouts.compute_returns(gamma, gae_lambda)                                    # tools/learn.py:212
    p_out = agent.update(rollouts, expert_loader if bcgail else None)               # tools/learn.py:218-223
    rollouts.after_update()                                                         # tools/learn.py:269
    return (d_out, p_out, *extra)
