"""Gradient-level parity helpers: one PPO minibatch and one WDGAIL critic step through the product's engines, gradients
read out of the flat gradient buffer *before* clipping / Adam, against torch autograd on the oracle restatement
(oracle/ref_path.py: algo/ppo.py:64-114, algo/wdgail.py:112-139) with identical parameters and inputs.

Used by tests/test_grads_cpu.py (host logic on the emulated ABI, fp32 re-association only) and tests/test_grads_gpu.py
(the CUDA kernels, TF32 contractions)."""
from types import SimpleNamespace as NS

import numpy as np
import torch

HP = dict(clip_param=0.1, value_loss_coef=0.5, logstd=[-1.4, -3.2])


def _batch(B, seed):
    from gail_carla_b200 import synthetic
    g = torch.Generator().manual_seed(seed)
    return (synthetic.synth_obs(B, g), synthetic.synth_metrics(B, g), synthetic.synth_actions(B, g, 0.2))


def policy_grads(device, B, Be=0, seed=3, clipped=True):
    """-> (got, ref): {param name: gradient} of value_coef*value_loss + action_loss (+ BC mix when Be > 0)."""
    import gail_carla_b200 as G
    from gail_carla_b200 import _abi as A, synthetic
    from oracle import ref_path as O
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    torch.manual_seed(seed)
    pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False)
    params = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    pol.to(device)
    obs, met, act = _batch(B, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    olp = torch.randn(B, generator=g) * 0.5 - 1.0
    vold = torch.randn(B, generator=g) * 0.3
    ret = vold + torch.randn(B, generator=g) * 0.5
    adv = torch.randn(B, generator=g)
    gamma = 0.3
    w_act = (1.0 - gamma) if Be else 1.0
    eng = pol.engine
    eng.sync_params()
    ws = eng.workspace(B + Be)
    to = lambda t: t.to(device).contiguous()
    eng.load_inputs(to(obs), to(met), None, B)
    if Be:
        e_obs, e_met, e_act = _batch(Be, seed + 5)
        eng.load_inputs(to(e_obs), to(e_met), None, Be, row0=B)
    head = eng.forward(B + Be, training=True)
    d_head = ws.buf("dhead", ws.rows, 4)
    acc = torch.zeros(4, dtype=torch.float64, device=device)
    A.ppo_loss(head, to(act), to(olp), to(vold), to(ret), to(adv), None, d_head, None, None, acc, B, HP["logstd"], True,
               HP["clip_param"], HP["value_loss_coef"], w_act, 0, clipped_value=clipped)
    if Be:
        A.ppo_loss(head[B:], to(e_act), None, None, None, None, None, d_head[B:], None, None, acc, Be, HP["logstd"], True,
                   0.0, 0.0, gamma, 1)
    eng.backward(B + Be, d_head)
    got = {k: eng.flat.g(k).detach().cpu().clone() for k in params}

    leaf = O._leaf(params)
    values, logp, _, _, _ = O.evaluate_actions(leaf, obs, met, act, True, HP["logstd"])
    if clipped:
        vl, al = O.ppo_losses(values, logp, olp.view(-1, 1), adv.view(-1, 1), vold.view(-1, 1), ret.view(-1, 1), HP["clip_param"])
    else:   # algo/ppo.py:112-113
        _, al = O.ppo_losses(values, logp, olp.view(-1, 1), adv.view(-1, 1), vold.view(-1, 1), ret.view(-1, 1), HP["clip_param"])
        vl = 0.5 * (ret.view(-1, 1) - values).pow(2).mean()
    if Be:
        _, e_logp, _, _, _ = O.evaluate_actions(leaf, e_obs, e_met, e_act, True, HP["logstd"])
        al = gamma * (-e_logp.mean()) + (1 - gamma) * al
    (vl * HP["value_loss_coef"] + al).backward()
    ref = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return got, ref


def critic_grads(device, B, seed=4):
    """-> (got, ref, got_scalars, ref_scalars) for one Discriminator minibatch: loss = -(E tanh D_e - E tanh D_p) + gp."""
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    from oracle import ref_path as O
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    torch.manual_seed(seed)
    disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, device, 2.5e-4, 1e-8, (0.9, 0.99), 0.5)
    params = {k: v.detach().clone() for k, v in disc.state_dict().items()}
    disc.to(device)
    e = _batch(B, seed + 1)
    p = _batch(B, seed + 2)
    alpha = torch.rand(B, 1, 1, 1, generator=torch.Generator().manual_seed(seed + 3))
    eng = disc.engine
    eng.sync_params()
    eng.workspace(3 * B)
    to = lambda t: t.to(device).contiguous()
    eng.load_inputs(to(e[0]), to(e[1]), to(e[2]), None, B, 0)
    eng.load_inputs(to(p[0]), to(p[1]), to(p[2]), None, B, B)
    acc = torch.zeros(8, dtype=torch.float64, device=device)
    eng.update_step(B, to(alpha.view(B)), acc)
    got = {k: eng.flat.g(k).detach().cpu().clone() for k in params}
    s = acc.cpu().tolist()
    got_s = dict(wd=(s[2] - s[3]) / B, gp=10.0 * s[4] / B)

    leaf = O._leaf(params)
    pd, ed = O.disc_forward(leaf, *p), O.disc_forward(leaf, *e)
    wd = torch.tanh(ed).mean() - torch.tanh(pd).mean()
    gp = O.grad_penalty(leaf, e, p, alpha)
    (-wd + gp).backward()
    ref = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return got, ref, got_s, dict(wd=wd.item(), gp=gp.item())


class linearised:
    """LeakyReLU slope 1.0 on both sides (product engines and oracle): the networks become linear maps, no activation
    mask can flip, and what is left of the deviation is the arithmetic of the contractions themselves."""

    def __enter__(self):
        from gail_carla_b200 import engine as E
        from oracle import ref_path as O
        self.E, self.O, self.old = E, O, (E.SLOPE, O.LRELU)
        E.SLOPE, O.LRELU = 1.0, 1.0
        return self

    def __exit__(self, *exc):
        self.E.SLOPE, self.O.LRELU = self.old
        return False


def stock_tf32_policy_grads(B, seed=3):
    """The oracle's autograd gradients of the same PPO minibatch computed twice by stock PyTorch: on the CPU in fp32 and
    on the CUDA device with cuDNN / cuBLAS allowed to use TF32 (the reference's own GPU numerics).  -> (tf32, fp32)."""
    from oracle import ref_path as O
    torch.manual_seed(seed)
    params = O.init_policy_params()
    obs, met, act = _batch(B, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    olp = torch.randn(B, generator=g) * 0.5 - 1.0
    vold = torch.randn(B, generator=g) * 0.3
    ret = vold + torch.randn(B, generator=g) * 0.5
    adv = torch.randn(B, generator=g)
    out = []
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for dev in ("cuda", "cpu"):
            leaf = O._leaf({k: v.to(dev) for k, v in params.items()})
            to = lambda t: t.to(dev)
            values, logp, _, _, _ = O.evaluate_actions(leaf, to(obs), to(met), to(act), True, HP["logstd"])
            vl, al = O.ppo_losses(values, logp, to(olp).view(-1, 1), to(adv).view(-1, 1), to(vold).view(-1, 1), to(ret).view(-1, 1),
                                  HP["clip_param"])
            (vl * HP["value_loss_coef"] + al).backward()
            out.append({k: v.grad.detach().cpu() for k, v in leaf.items()})
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out[0], out[1]


def rel_errors(got, ref):
    """{name: (cosine, rel-Frobenius)} for the tensors whose reference gradient is non-zero."""
    out = {}
    for k, r in ref.items():
        g = got[k].double().reshape(-1); r = r.double().reshape(-1)
        if r.norm() == 0:
            continue
        out[k] = (float(torch.dot(g, r) / (r.norm() * g.norm() + 1e-300)), float((g - r).norm() / r.norm()))
    return out


def compare(got, ref, min_cos, max_rel, what):
    """Per tensor: cosine >= min_cos and ||got-ref||_F <= max_rel * ||ref||_F (tensors whose reference gradient is exactly
    zero must be zero).  Returns the worst (cos, rel) seen, for the test's printout."""
    worst_cos, worst_rel = 1.0, 0.0
    for k, r in ref.items():
        g = got[k].double().reshape(-1); r = r.double().reshape(-1)
        nr, ng = r.norm().item(), g.norm().item()
        if nr == 0.0:
            assert ng == 0.0, f"{what} {k}: reference gradient is exactly 0, got norm {ng:.3g}"
            continue
        cos = float(torch.dot(g, r) / (nr * ng + 1e-300))
        rel = float((g - r).norm() / nr)
        worst_cos, worst_rel = min(worst_cos, cos), max(worst_rel, rel)
        assert np.isfinite(rel) and cos >= min_cos and rel <= max_rel, \
            f"{what} {k}: cosine {cos:.6f} (min {min_cos}), rel-Frobenius {rel:.3e} (max {max_rel})"
    return worst_cos, worst_rel
