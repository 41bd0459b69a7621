"""Golden scalar log of the UNMODIFIED reference training loop (tools/learn.py::gailLearning_mujoco_origin) on synthetic envs.

Run in the build container only:  python tests/golden/make_learn_golden.py

`tools/learn.py` cannot be imported as is here (it needs tensorboardX, and tools/envs.py needs the CARLA client), so the
three missing third-party modules are replaced by stand-ins *in sys.modules* before the import - the reference files
themselves are executed unmodified from /root/reference:
  * ``tensorboardX.SummaryWriter``  -> a recorder that keeps every ``add_scalar(title, value, step)`` call,
  * ``carla_env`` (CARLA client wrapper) and ``gym`` -> empty stand-ins (only imported, never used by the loop).
The loop then runs with the reference's own RolloutStorage / Policy / PPO / Discriminator on the CPU against
``gail_carla_b200.synthetic.SyntheticVecEnv`` / ``SyntheticEvalEnv`` (the vec-env protocol of tools/envs.py) and the scalar
stream is stored in tests/golden/learn_loop.json for tests/test_learn_cpu.py / test_learn_gpu.py.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types
from types import SimpleNamespace as NS

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

RECORD = []


class _Writer:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, title, value, step):
        RECORD.append([str(title), None if value is None else float(value), int(step)])


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_stub("tensorboardX", SummaryWriter=_Writer)
_stub("carla_env", CarlaEnv=object)
if "gym" not in sys.modules:
    try:
        import gym  # noqa: F401
    except Exception:
        _stub("gym", Env=object, spaces=types.SimpleNamespace(Box=object))

from gail_carla_b200 import synthetic  # noqa: E402
from tools import learn as ref_learn, utli as ref_utli  # noqa: E402
from tools.model import Policy as RefPolicy  # noqa: E402
from algo.ppo import PPO as RefPPO  # noqa: E402
from algo.wdgail import Discriminator as RefDisc  # noqa: E402

HP = dict(lr=1e-4, eps=1e-8, betas=(0.9, 0.99), clip_param=0.1, value_loss_coef=0.5, max_grad_norm=0.5,
          gail_lr=2.5e-4, gail_eps=1e-8, gail_betas=(0.9, 0.99), gail_max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95,
          logstd=[-1.4, -3.2])
CASE = dict(nenv=4, nsteps=32, updates=2, B=16, n_train=2, n_val=1, mean_episode_len=4, eval_len=6)


def run_params(case):
    return dict(num_steps=case["nsteps"], num_env_steps=case["updates"] * case["nsteps"], envs_params=[{}] * case["nenv"], routes=[0],
                lr=HP["lr"], use_linear_lr_decay=True, gail_epoch=1, gail_pre_epoch=1, gail_thre=0, gamma=HP["gamma"],
                gae_lambda=HP["gae_lambda"], bcgail=False, eval_interval=1, log_interval=1, resume_training=False,
                algo="ppo", env_name="synthetic", seed=1, gail_batch_size=case["B"])


def main():
    torch.set_num_threads(os.cpu_count())
    c = CASE
    torch.manual_seed(1); np.random.seed(1)
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    pol = RefPolicy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False)
    agent = RefPPO(pol, HP["clip_param"], 1, c["B"], HP["value_loss_coef"], "cpu", lr=HP["lr"], eps=HP["eps"], betas=HP["betas"],
                   max_grad_norm=HP["max_grad_norm"], gamma=None, decay=None, act_space=asp)
    disc = RefDisc(synthetic.OBS_SHAPE, sp, asp, 100, "cpu", HP["gail_lr"], HP["gail_eps"], HP["gail_betas"], HP["gail_max_grad_norm"])
    envs = synthetic.SyntheticVecEnv(c["nenv"], seed=3, device="cpu", mean_episode_len=c["mean_episode_len"], routes=(0,))
    env_eval = synthetic.SyntheticEvalEnv(ep_length=c["eval_len"], seed=5, device="cpu")
    train = synthetic.SyntheticExpertLoader(c["n_train"], c["B"], seed=21)
    val = synthetic.SyntheticExpertLoader(c["n_val"], c["B"], seed=22)
    torch.manual_seed(7)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:      # the loop writes ./runs/... and ./gail_model.pt
        os.chdir(tmp)
        try:
            ref_learn.gailLearning_mujoco_origin(run_params(c), envs, env_eval, pol, agent, disc, train, val, "cpu", ref_utli)
            ckpt = torch.load("gail_model.pt")
        finally:
            os.chdir(cwd)
    out = dict(case=c, scalars=RECORD, checkpoint_update=int(ckpt[2]),
               policy_param_sum=float(sum(v.double().sum() for v in ckpt[0].values())),
               disc_param_sum=float(sum(v.double().sum() for v in ckpt[1].values())))
    with open(os.path.join(HERE, "learn_loop.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print(f"{len(RECORD)} scalars recorded; titles: {sorted(set(r[0] for r in RECORD))}")


if __name__ == "__main__":
    main()
