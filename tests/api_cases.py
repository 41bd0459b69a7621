"""Reference-API edge cases shared by the CPU (emulated ABI) and GPU test modules.  Every check compares the drop-in
classes with torch autograd / torch CPU on the oracle restatement (oracle/ref_path.py) at identical parameters."""
from types import SimpleNamespace as NS

import numpy as np
import torch

import grad_cases as GC

LOGSTD = [-1.4, -3.2]


def _policy(device, seed=5):
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    torch.manual_seed(seed)
    pol = G.Policy(synthetic.OBS_SHAPE, NS(shape=(4,)), NS(shape=(2,)), True, LOGSTD, False)
    params = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    return pol.to(device), params


def _critic(device, seed=6):
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    torch.manual_seed(seed)
    disc = G.Discriminator(synthetic.OBS_SHAPE, NS(shape=(4,)), NS(shape=(2,)), 100, device, 2.5e-4, 1e-8, (0.9, 0.99), 0.5)
    params = {k: v.detach().clone() for k, v in disc.state_dict().items()}
    return disc.to(device), params


def evaluate_actions_is_differentiable(device, B, min_cos, max_rel):
    """tools/model.py:45-53 as the reference's learn_bc.py:37-45 uses it: loss built from the returned log-probs and
    values, ``loss.backward()``, gradients in ``p.grad``; a second backward accumulates."""
    from oracle import ref_path as O
    pol, params = _policy(device)
    obs, met, act = GC._batch(B, 31)
    w = torch.randn(B, 1, generator=torch.Generator().manual_seed(3))
    value, logp, ent, s0, s1 = pol.evaluate_actions(obs.to(device), met.to(device), act.to(device))
    assert value.requires_grad and logp.requires_grad
    loss = -(logp.mean()) + 0.25 * (value * w.to(device)).sum()
    loss.backward()
    got = {k: p.grad.detach().cpu().clone() for k, p in pol.named_parameters()}
    leaf = O._leaf(params)
    v_r, lp_r, ent_r, _, _ = O.evaluate_actions(leaf, obs, met, act, True, LOGSTD)
    (-(lp_r.mean()) + 0.25 * (v_r * w).sum()).backward()
    ref = {k: v.grad for k, v in leaf.items()}
    GC.compare(got, ref, min_cos, max_rel, "evaluate_actions")
    fwd_tol = min(max_rel, 5e-3)       # forward values: no mask-flip amplification, TF32 level at most
    assert torch.allclose(value.detach().cpu(), v_r.detach(), rtol=fwd_tol * 10, atol=fwd_tol)
    assert torch.allclose(logp.detach().cpu(), lp_r.detach(), rtol=fwd_tol * 10, atol=fwd_tol * 10)
    assert abs(float(ent) - float(ent_r)) < 1e-6
    # accumulation semantics of autograd: a second identical pass doubles p.grad
    value, logp, *_ = pol.evaluate_actions(obs.to(device), met.to(device), act.to(device))
    (-(logp.mean()) + 0.25 * (value * w.to(device)).sum()).backward()
    got2 = {k: p.grad.detach().cpu() for k, p in pol.named_parameters()}
    GC.compare(got2, {k: 2 * v for k, v in ref.items()}, min_cos, max_rel, "evaluate_actions (accumulated)")
    # no-grad call returns plain tensors
    with torch.no_grad():
        v2, lp2, *_ = pol.evaluate_actions(obs.to(device), met.to(device), act.to(device))
    assert not v2.requires_grad and torch.allclose(v2.cpu(), value.detach().cpu(), atol=1e-6)


def forward_gp_returns_first_order_handles(device, B, tol):
    """algo/wdgail.py:40-54 with gp=True: 4-tuple, and autograd.grad(output, state_transformed, ones) = dD/dx."""
    from oracle import ref_path as O
    disc, params = _critic(device)
    obs, met, act = GC._batch(B, 41)
    out, st, mt, at = disc(obs.to(device), met.to(device), act.to(device), gp=True)
    assert out.shape == (B, 1) and st.shape == obs.shape and mt.shape == (B, 13) and at.shape == (B, 2)
    assert st.requires_grad and st.is_leaf
    g = torch.autograd.grad(out, st, torch.ones_like(out))[0].cpu()
    x = obs.clone().requires_grad_(True)
    d = O.disc_forward(params, x, met, act)
    g_ref = torch.autograd.grad(d, x, torch.ones_like(d))[0]
    fwd_tol = min(tol, 5e-3)
    assert torch.allclose(out.detach().cpu(), d.detach(), rtol=fwd_tol * 10, atol=fwd_tol)
    rel = float((g - g_ref).norm() / g_ref.norm())
    print(f"forward(gp=True): dD/dx rel-Frobenius {rel:.3e}")
    assert rel <= tol, f"dD/dx rel-Frobenius {rel:.3e}"
    assert torch.allclose(mt.cpu(), O.metrics_features(params["metrics_processor.road_option_embedding.weight"], met), rtol=1e-5, atol=1e-5)
    assert torch.equal(st.detach().cpu(), obs) and torch.equal(at.cpu(), act)


def expert_loader_shorter_and_dropped_remainder(device, tol):
    """zip(expert_loader, generator) stops at the shorter side (algo/wdgail.py:112): 2 expert batches against 5 policy
    minibatches -> 2 optimiser steps; compute_loss(batch_size=...) permutes only the first `batch_size` rows and drops
    the partial last minibatch (algo/wdgail.py:158, tools/storage.py:57-63)."""
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    from oracle import ref_path as O
    disc, params = _critic(device, seed=8)
    T, N, B = 11, 2, 4
    ro = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device="cpu")
    synthetic.fill_rollout(ro, seed=77)
    o_ro = {k: getattr(ro, k).clone() for k in ("obs", "metrics", "actions", "gail_rewards")}
    if device != "cpu":
        ro_d = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device=device)
        for k in ("obs", "metrics", "actions"):
            getattr(ro_d, k).copy_(getattr(ro, k))
        ro = ro_d
    loader = synthetic.SyntheticExpertLoader(2, B, seed=78)
    adam = O.AdamState(params, 2.5e-4, 1e-8, (0.9, 0.99))
    torch.manual_seed(9)
    ref_cl = O.disc_compute_loss(params, loader, o_ro, batch_size=10)        # 10 rows -> 2 full minibatches of 4
    ref_up = O.disc_update(params, adam, loader, o_ro, 0.5)                  # 22 rows -> 5 minibatches, 2 expert batches
    torch.manual_seed(9)
    got_cl = disc.compute_loss(loader, ro, batch_size=10)
    got_up = disc.update(loader, ro)
    assert disc.optimizer.t == 2, "two expert batches -> two optimiser steps"
    np.testing.assert_allclose(np.array(got_cl), np.array(ref_cl), rtol=tol, atol=tol * 0.1)
    np.testing.assert_allclose(np.array(got_up), np.array(ref_up), rtol=tol, atol=tol * 0.1)
    # an empty pairing (batch_size smaller than one minibatch) returns the reference's (0, 0, 0)
    assert tuple(disc.compute_loss(loader, ro, batch_size=3)) == (0, 0, 0)


def reward_saturates_to_inf(device):
    """algo/wdgail.py:185-186: -log(1 - sigmoid(d)) overflows to +inf once sigmoid(d) rounds to 1.0f (d >~ 16.6 in fp32).
    The tail is part of the reference's behaviour: finite and close below the threshold, +inf above it."""
    from gail_carla_b200 import _abi as A
    d = torch.tensor([-30.0, -5.0, 0.0, 5.0, 12.0, 15.0, 16.0, 17.5, 18.0, 25.0, 88.0, 200.0])
    ref = -(1 - torch.sigmoid(d)).log()
    out = torch.empty_like(d).to(device)
    A.reward_epilogue(d.to(device), out, d.numel())
    out = out.cpu()
    assert torch.isinf(ref[7:]).all() and torch.isfinite(ref[:7]).all()          # the reference's own behaviour
    assert torch.isinf(out[7:]).all() and (out[7:] > 0).all(), out
    assert torch.allclose(out[:6], ref[:6], rtol=2e-3, atol=1e-7), (out, ref)
    assert torch.isfinite(out[6]) and abs(float(out[6]) - float(ref[6])) <= 0.2   # 1-sigmoid(16) is 1-2 ulp of 1.0
