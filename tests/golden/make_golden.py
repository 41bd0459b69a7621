"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case it (1) builds the reference modules at seed 1 (wdail_carla.py:152-154,200-237),
(2) fills a reference ``RolloutStorage`` with ``gail_carla_b200.synthetic`` data, (3) replays the
env-free slice of ``tools/learn.py:137-223,269`` on the reference classes, (4) replays the same slice
on ``oracle.ref_path`` and asserts the two agree, and (5) stores the *reference's* outputs.
"""
from __future__ import annotations

import copy
import os
import sys
import time
from types import SimpleNamespace as NS

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from gail_carla_b200 import synthetic  # noqa: E402
from oracle import ref_path as O  # noqa: E402

from tools.storage import RolloutStorage as RefStorage  # noqa: E402
from tools.model import Policy as RefPolicy  # noqa: E402
from algo.ppo import PPO as RefPPO  # noqa: E402
from algo.wdgail import Discriminator as RefDisc  # noqa: E402
from common.running_mean_std import RunningMeanStd as RefRMS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
HP = dict(lr=1e-4, eps=1e-8, betas=(0.9, 0.99), clip_param=0.1, value_loss_coef=0.5, max_grad_norm=0.5,
          gail_lr=2.5e-4, gail_eps=1e-8, gail_betas=(0.9, 0.99), gail_max_grad_norm=0.5,
          gamma=0.99, gae_lambda=0.95, logstd=[-1.4, -3.2])


def param_digest(sd):
    """Small fingerprint of a state_dict: full tensor when tiny, else sums + a strided sample."""
    out = {}
    for k, v in sd.items():
        v = v.detach().float().reshape(-1)
        if v.numel() <= 4096:
            out[k + "|full"] = v.numpy().copy()
        else:
            stride = v.numel() // 2048
            out[k + "|sample"] = v[::stride][:2048].numpy().copy()
        out[k + "|sum"] = np.array([v.double().sum().item(), v.double().abs().sum().item()])
    return out


def storage_dict(ro):
    return {k: getattr(ro, k) for k in ("obs", "metrics", "actions", "action_log_probs", "value_preds",
                                        "returns", "masks", "gail_rewards", "rewards")}


def gae_case(T, N, seed):
    ro = RefStorage(T, N, (1, 2, 2), (4,), (2,))
    g = torch.Generator().manual_seed(seed)
    ro.gail_rewards.copy_(torch.nn.functional.softplus(torch.randn(T, N, 1, generator=g)))
    ro.value_preds.copy_(torch.randn(T + 1, N, 1, generator=g))
    m = (torch.rand(T + 1, N, 1, generator=g) >= 0.02).float()
    ro.masks.copy_(m)
    ro.compute_returns(HP["gamma"], HP["gae_lambda"])
    adv = ro.returns[:-1] - ro.value_preds[:-1]
    adv_n = (adv - adv.mean()) / (adv.std() + 1e-5)
    mine = O.gae_returns(ro.gail_rewards, ro.value_preds, ro.masks, HP["gamma"], HP["gae_lambda"])
    assert torch.equal(mine, ro.returns), "oracle GAE differs from reference"
    assert torch.equal(O.normalized_advantages(ro.returns, ro.value_preds), adv_n)
    return dict(gail_rewards=ro.gail_rewards.numpy(), value_preds=ro.value_preds.numpy(), masks=ro.masks.numpy(),
                returns=ro.returns.numpy(), adv_norm=adv_n.numpy(), gamma=HP["gamma"], gae_lambda=HP["gae_lambda"])


def update_case(T, N, B_ppo, B_gail, ppo_epoch, gail_epoch, n_expert, bc, seed=1, clipped=True):
    t0 = time.time()
    sp = NS(shape=(4,)); asp = NS(shape=(2,))
    torch.manual_seed(seed); np.random.seed(seed)
    pol = RefPolicy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False)
    agent = RefPPO(pol, HP["clip_param"], ppo_epoch, B_ppo, HP["value_loss_coef"], "cpu", lr=HP["lr"], eps=HP["eps"],
                   betas=HP["betas"], max_grad_norm=HP["max_grad_norm"], gamma=0.3 if bc else None,
                   decay=0.9 if bc else None, act_space=asp, use_clipped_value_loss=clipped)
    disc = RefDisc(synthetic.OBS_SHAPE, sp, asp, 100, "cpu", HP["gail_lr"], HP["gail_eps"], HP["gail_betas"],
                   HP["gail_max_grad_norm"])
    # oracle twins built from the same seed must be identical to the reference modules
    torch.manual_seed(seed)
    o_pol = O.init_policy_params(); o_disc = O.init_disc_params()
    for k, v in pol.state_dict().items():
        assert torch.equal(v, o_pol[k]), k
    for k, v in disc.state_dict().items():
        assert torch.equal(v, o_disc[k]), k
    o_padam = O.AdamState(o_pol, HP["lr"], HP["eps"], HP["betas"])
    o_dadam = O.AdamState(o_disc, HP["gail_lr"], HP["gail_eps"], HP["gail_betas"])

    ro = RefStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,))
    synthetic.fill_rollout(ro, seed=seed + 10)
    o_ro = {k: v.clone() for k, v in storage_dict(ro).items()}
    loader = synthetic.SyntheticExpertLoader(n_expert, B_gail, seed=seed + 20)
    out = {}

    # ---- reference: tools/learn.py:137-223,269 ----
    torch.manual_seed(seed + 100)
    with torch.no_grad():
        ro.value_preds[-1] = pol.get_value(ro.obs[-1], ro.metrics[-1]).detach()
    out["bootstrap_value"] = ro.value_preds[-1].numpy().copy()
    out["compute_loss_before"] = np.array(disc.compute_loss(loader, ro))
    d_tuples = [disc.update(loader, ro) for _ in range(gail_epoch)]
    out["disc_update"] = np.array(d_tuples)
    out["compute_loss_after"] = np.array(disc.compute_loss(loader, ro))
    for step in range(T):
        ro.gail_rewards[step] = disc.predict_reward(ro.obs[step], ro.metrics[step], ro.actions[step],
                                                    HP["gamma"], ro.masks[step])
    out["gail_rewards"] = ro.gail_rewards.numpy().copy()
    ro.compute_returns(HP["gamma"], HP["gae_lambda"])
    out["returns"] = ro.returns.numpy().copy()
    p_tuple = agent.update(ro, loader if bc else None)
    out["ppo_update"] = np.array([np.nan if x is None else float(x) for x in p_tuple])
    ro.after_update()
    with torch.no_grad():
        v, a, lp = pol.act(ro.obs[:4, 0], ro.metrics[:4, 0], deterministic=True)
    out["act_value"], out["act_action"], out["act_logp"] = v.numpy(), a.numpy(), lp.numpy()
    for k, v in param_digest(pol.state_dict()).items():
        out["pol|" + k] = v
    for k, v in param_digest(disc.state_dict()).items():
        out["disc|" + k] = v

    # ---- oracle restatement on the same inputs / same default-generator stream ----
    torch.manual_seed(seed + 100)
    with torch.no_grad():
        o_ro["value_preds"][-1] = O.policy_base(o_pol, o_ro["obs"][-1], o_ro["metrics"][-1], True, HP["logstd"])[0]
    o_cl0 = O.disc_compute_loss(o_disc, loader, o_ro)
    o_d = [O.disc_update(o_disc, o_dadam, loader, o_ro, HP["gail_max_grad_norm"]) for _ in range(gail_epoch)]
    o_cl1 = O.disc_compute_loss(o_disc, loader, o_ro)
    for step in range(T):
        o_ro["gail_rewards"][step] = O.predict_reward(o_disc, o_ro["obs"][step], o_ro["metrics"][step],
                                                      o_ro["actions"][step])
    o_ro["returns"] = O.gae_returns(o_ro["gail_rewards"], o_ro["value_preds"], o_ro["masks"], HP["gamma"],
                                    HP["gae_lambda"])
    o_p = O.ppo_update(o_pol, o_padam, o_ro, clip_param=HP["clip_param"], ppo_epoch=ppo_epoch,
                       mini_batch_size=B_ppo, value_loss_coef=HP["value_loss_coef"],
                       max_grad_norm=HP["max_grad_norm"], logstd=HP["logstd"],
                       expert_loader=loader if bc else None, bc_gamma=0.3 if bc else None,
                       decay=0.9 if bc else None, use_clipped_value_loss=clipped)

    def close(a, b, what, rtol=2e-4, atol=2e-5):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        ok = np.allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)
        err = np.nanmax(np.abs(a - b) / (np.abs(b) + atol / rtol)) if a.size else 0
        print(f"   oracle vs reference  {what:24s} max-rel {err:.2e} {'ok' if ok else 'MISMATCH'}")
        assert ok, what

    close(o_ro["value_preds"][-1].numpy(), out["bootstrap_value"], "bootstrap value")
    close(o_cl0, out["compute_loss_before"], "compute_loss before")
    close(o_d, out["disc_update"], "disc.update tuple", rtol=1e-3)
    close(o_cl1, out["compute_loss_after"], "compute_loss after", rtol=1e-3, atol=1e-4)
    close(o_ro["gail_rewards"].numpy(), out["gail_rewards"], "gail_rewards", rtol=1e-3, atol=1e-4)
    close(o_ro["returns"].numpy(), out["returns"], "returns", rtol=1e-3, atol=1e-4)
    close([np.nan if x is None else x for x in o_p], out["ppo_update"], "ppo.update tuple", rtol=1e-3, atol=1e-4)
    for k, v in param_digest(o_pol).items():
        close(v, out["pol|" + k], "pol " + k[-40:], rtol=1e-3, atol=3e-4 if "sum" not in k else 1e-1)
    for k, v in param_digest(o_disc).items():
        close(v, out["disc|" + k], "disc " + k[-40:], rtol=1e-3, atol=6e-4 if "sum" not in k else 1e-1)
    out["config"] = np.array([T, N, B_ppo, B_gail, ppo_epoch, gail_epoch, n_expert, int(bc), seed, int(clipped)])
    print(f"   case done in {time.time() - t0:.1f}s")
    return out


def rms_case():
    rms = RefRMS(shape=())
    g = np.random.RandomState(3)
    xs = [g.randn(n) * s + m for n, s, m in ((17, 1.0, 0.0), (5, 3.0, 2.0), (256, 0.1, -4.0))]
    hist = []
    st = (np.zeros(()), np.ones(()), 1e-4)
    for x in xs:
        rms.update(x)
        st = O.rms_update(st, x)
        assert np.allclose(st[0], rms.mean) and np.allclose(st[1], rms.var) and np.isclose(st[2], rms.count)
        hist.append([rms.mean, rms.var, rms.count])
    return dict(x0=xs[0], x1=xs[1], x2=xs[2], hist=np.array(hist, dtype=np.float64))


def main():
    torch.set_num_threads(os.cpu_count())
    if "--only-new" in sys.argv:      # round 2 additions (the older files are reproducible with the default run)
        print("update unclipped (T=6,N=2,B=6, use_clipped_value_loss=False)")
        np.savez_compressed(os.path.join(HERE, "update_unclipped.npz"),
                            **update_case(T=6, N=2, B_ppo=6, B_gail=6, ppo_epoch=1, gail_epoch=1, n_expert=2, bc=False, clipped=False))
        print("update mid (T=64,N=8,B=256: 2 minibatches per epoch, BC mix on, multi-tile / split-K shapes)")
        np.savez_compressed(os.path.join(HERE, "update_mid.npz"),
                            **update_case(T=64, N=8, B_ppo=256, B_gail=256, ppo_epoch=1, gail_epoch=1, n_expert=2, bc=True))
        return
    print("gae cases")
    np.savez_compressed(os.path.join(HERE, "gae_64x4.npz"), **gae_case(64, 4, 5))
    np.savez_compressed(os.path.join(HERE, "gae_2048x16.npz"), **gae_case(2048, 16, 6))
    np.savez_compressed(os.path.join(HERE, "gae_33x1.npz"), **gae_case(33, 1, 7))
    np.savez_compressed(os.path.join(HERE, "rms.npz"), **rms_case())
    print("update tiny (T=8,N=2,B=8, 2 ppo epochs, BC mix on)")
    np.savez_compressed(os.path.join(HERE, "update_tiny.npz"),
                        **update_case(T=8, N=2, B_ppo=8, B_gail=8, ppo_epoch=2, gail_epoch=1, n_expert=2, bc=True))
    print("update tiny2 (T=6,N=4,B=8, no BC, 2 gail epochs)")
    np.savez_compressed(os.path.join(HERE, "update_tiny2.npz"),
                        **update_case(T=6, N=4, B_ppo=12, B_gail=8, ppo_epoch=1, gail_epoch=2, n_expert=3, bc=False))
    if "--c1" in sys.argv:
        print("config 1 (T=128,N=1,B=128)")
        np.savez_compressed(os.path.join(HERE, "update_c1.npz"),
                            **update_case(T=128, N=1, B_ppo=128, B_gail=128, ppo_epoch=1, gail_epoch=1, n_expert=1,
                                          bc=True))


if __name__ == "__main__":
    main()
