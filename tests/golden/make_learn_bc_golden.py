"""Golden loss series of the UNMODIFIED reference behaviour-cloning loop (learn_bc.py::learn_bc, lines 15-72).

Run in the build container only:  python tests/golden/make_learn_bc_golden.py

learn_bc.py imports gym, the CARLA client wrapper and tensorboardX at module level; the three are replaced by stand-ins in
sys.modules (the reference file itself is executed unmodified).  Its loop runs a hard-coded 300 epochs; the stand-in
SummaryWriter records ``add_scalar`` and stops the loop after EPOCHS epochs by raising a private exception.
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
EPOCHS = 3
RECORD = []


class _Stop(Exception):
    pass


class _Writer:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, title, value, step):
        RECORD.append([str(title), float(value), int(step)])
        if len(RECORD) >= 2 * EPOCHS:
            raise _Stop()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


_stub("tensorboardX", SummaryWriter=_Writer)
_stub("carla_env", CarlaEnv=object)
try:
    import gym  # noqa: F401
except Exception:
    _stub("gym", spaces=types.SimpleNamespace(Box=object))

from gail_carla_b200 import synthetic  # noqa: E402
import learn_bc as ref_bc  # noqa: E402   (the reference's /root/reference/learn_bc.py)
from tools.model import Policy as RefPolicy  # noqa: E402

LOGSTD = [-1.4, -3.2]
CASE = dict(B=8, n_train=2, B_eval=6, n_eval=1, seed=1)


def main():
    torch.set_num_threads(os.cpu_count())
    c = CASE
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    pol = RefPolicy(synthetic.OBS_SHAPE, NS(shape=(4,)), NS(shape=(2,)), True, LOGSTD, False)
    train = synthetic.SyntheticExpertLoader(c["n_train"], c["B"], seed=31)
    val = synthetic.SyntheticExpertLoader(c["n_eval"], c["B_eval"], seed=32)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ref_bc.learn_bc(pol, "cpu", train, val)
        except _Stop:
            pass
        finally:
            os.chdir(cwd)
    out = dict(case=c, epochs=EPOCHS, scalars=RECORD)
    with open(os.path.join(HERE, "learn_bc.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print(RECORD)


if __name__ == "__main__":
    main()
