"""CPU tests of the training-iteration orchestration (gail_carla_b200/learn.py) with the
C-ABI replaced by its CPU statements (`emulated_abi`): schedules and bookkeeping against restatements of the reference
lines they follow, scalar titles against tools/utli.py, checkpoint format round trip, and a 2-iteration run of the
whole loop on the synthetic vec-env."""
import math
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

HP = dict(lr=1e-4, eps=1e-8, betas=(0.9, 0.99), clip_param=0.1, value_loss_coef=0.5, max_grad_norm=0.5,
          gail_lr=2.5e-4, gail_eps=1e-8, gail_betas=(0.9, 0.99), gail_max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95,
          logstd=[-1.4, -3.2])


def test_schedules_follow_the_reference_lines():
    from gail_carla_b200 import learn as L
    # tools/utli.py:121-125
    for ep, tot, lr in ((1, 10, 1e-4), (10, 10, 3e-4), (3, 7, 2.5e-4)):
        assert L.linear_lr(lr, ep, tot) == lr - (lr * (ep / float(tot)))
    # tools/learn.py:146-151
    for i_update in range(1, 12):
        ge, pre, thre = 1, 5, 8
        ref = ge
        if i_update < thre:
            ref += (pre - ge) * (thre - (i_update - 1)) / thre
            ref = int(ref)
        assert L.gail_epochs_for(i_update, ge, pre, thre) == ref
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    L.set_lr(opt, 0.25)
    assert all(g["lr"] == 0.25 for g in opt.param_groups)


def test_episodic_gail_returns_equals_the_reference_loop():
    from gail_carla_b200 import learn as L
    g = torch.Generator().manual_seed(5)
    T, N = 97, 5
    for trial in range(3):
        rewards = torch.rand(T, N, 1, generator=g)
        masks = (torch.rand(T + 1, N, 1, generator=g) > 0.08).float()
        carry0 = [float(v) for v in torch.rand(N, generator=g)]
        # tools/learn.py:196-209 restated
        cum, buf = list(carry0), []
        for step in range(T):
            for i_env in range(N):
                if masks[step][i_env]:
                    cum[i_env] += rewards[step][i_env].item()
                else:
                    buf.append(cum[i_env])
                    cum[i_env] = .0
        carry = list(carry0)
        got = L.episodic_gail_returns(rewards, masks, carry)
        assert len(got) == len(buf)
        np.testing.assert_allclose(got, buf, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(carry, cum, rtol=1e-6, atol=1e-7)


def test_scalar_titles_are_the_reference_titles():
    from gail_carla_b200 import learn as L
    # tools/utli.py:9-18, 35-49, 70-79
    assert L.PPO_SCALARS == ("ppo_value", "ppo_loss", "ppo_entropy", "bc_loss", "gail_loss", "gail_gamma", "steer_std", "throttle_std")
    assert len(L.DISC_SCALARS) == 13 and L.DISC_SCALARS[7:] == ("disc_pre_loss", "expert_pre_reward", "policy_pre_reward",
                                                              "disc_after_loss", "expert_after_reward", "policy_after_reward")
    assert L.TRAIN_SCALARS[:3] == ("Train reward", "Train steps", "Expert reward")
    seen = []
    log = L.ScalarLog(NS(add_scalar=lambda t, v, s: seen.append((t, s))))
    log.record_routes({3: [1.0, -2.0], 4: []}, 7)
    assert seen == [("route_03_max_reward", 7), ("route_03_min_reward", 7)]


def test_training_loop_runs_and_checkpoints(emulated_abi, tmp_path):
    import gail_carla_b200 as G
    from gail_carla_b200 import learn as L, synthetic
    torch.manual_seed(1); np.random.seed(1)
    nenv, nsteps = 2, 8                      # 4 steps per env per iteration
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False)
    agent = G.PPO(pol, HP["clip_param"], 1, 4, HP["value_loss_coef"], "cpu", lr=HP["lr"], eps=HP["eps"], betas=HP["betas"],
                  max_grad_norm=HP["max_grad_norm"], gamma=None, decay=None, act_space=asp)
    disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, "cpu", HP["gail_lr"], HP["gail_eps"], HP["gail_betas"],
                           HP["gail_max_grad_norm"])
    envs = synthetic.SyntheticVecEnv(nenv, seed=3, mean_episode_len=3, routes=(0, 1))
    env_eval = synthetic.SyntheticEvalEnv(ep_length=4, seed=5)
    train = synthetic.SyntheticExpertLoader(2, 4, seed=21)
    val = synthetic.SyntheticExpertLoader(1, 4, seed=22)
    rp = dict(num_steps=nsteps, num_env_steps=2 * nsteps, envs_params=[{}] * nenv, routes=[0, 1], lr=HP["lr"],
              use_linear_lr_decay=True, gail_epoch=1, gail_pre_epoch=2, gail_thre=2, gamma=HP["gamma"], gae_lambda=HP["gae_lambda"],
              bcgail=False, eval_interval=1, log_interval=1, resume_training=False)
    path = str(tmp_path / "gail_model.pt")
    before = {k: v.clone() for k, v in pol.state_dict().items()}
    log = L.gail_learning(rp, envs, env_eval, pol, agent, disc, train, val, "cpu", model_path=path)
    steps = {r["step"] for r in log.history}
    assert steps == {1, 2}
    rows = [r for r in log.history if "ppo_value" in r]
    assert len(rows) == 2 and all(math.isfinite(r["ppo_value"]) and math.isfinite(r["ppo_loss"]) for r in rows)
    drows = [r for r in log.history if "dis_total_loss" in r]
    assert len(drows) == 2 and all(math.isfinite(r["dis_gp"]) and math.isfinite(r["disc_after_loss"]) for r in drows)
    assert agent.optimizer.param_groups[0]["lr"] == L.linear_lr(HP["lr"], 2, 2)
    assert any((pol.state_dict()[k] != before[k]).any() for k in before)
    # checkpoint list format of tools/learn.py:290-291 and the resume path (:81-87)
    assert os.path.exists(path)
    data = torch.load(path, map_location="cpu")
    assert isinstance(data, list) and len(data) == 4 and data[2] == 2 and set(data[0]) == set(pol.state_dict())
    pol2 = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False)
    disc2 = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, "cpu", HP["gail_lr"], HP["gail_eps"], HP["gail_betas"],
                            HP["gail_max_grad_norm"])
    i_update, elapsed = L.load_checkpoint(path, pol2, disc2)
    assert i_update == 2 and elapsed >= 0
    for k, v in pol.state_dict().items():
        assert torch.equal(v.cpu(), pol2.state_dict()[k].cpu())
    # resuming a finished run does no further updates
    rp2 = dict(rp, resume_training=True)
    log2 = L.gail_learning(rp2, envs, env_eval, pol2, agent, disc2, train, val, "cpu", model_path=path)
    assert log2.history == []


def test_training_loop_matches_the_unmodified_reference_loop(emulated_abi, tmp_path):
    """tools/learn.py::gailLearning_mujoco_origin (run unmodified by tests/golden/make_learn_golden.py with stand-ins only for
    the tensorboardX / CARLA-client imports) vs gail_learning on the CPU statements of the ABI: the same scalar stream - titles,
    order, step numbers - with values at fp32 re-association level (2e-3; the env-driven scalars exactly)."""
    import learn_cases as LC
    gold, rows, ckpt = LC.run("cpu", tmp_path, "cpu")
    worst = LC.check(gold, rows, ckpt, tol=2e-3)
    print("worst relative deviation from the reference loop:", worst)


def test_training_loop_on_the_byte_store_matches_the_reference_loop(emulated_abi, tmp_path):
    """Same run with the rollout in the uint8 observation store (insert / act / predict_reward / update all on ByteObs):
    lossless, so the same golden at the same tolerance."""
    import learn_cases as LC
    gold, rows, ckpt = LC.run("cpu", tmp_path, "cpu_u8", obs_dtype=torch.uint8)
    LC.check(gold, rows, ckpt, tol=2e-3)
