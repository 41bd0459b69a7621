"""Behaviour-cloning pre-training (gail_carla_b200/learn_bc.py) on the CPU statements of the ABI against a restatement
of learn_bc.py:15-72 built from the oracle's policy forward (oracle/ref_path.py, pinned to the unmodified reference)
and torch autograd + Adam: same batches (device-resident loader vs DataLoader-order host batches), epoch losses within
1e-4 relative, parameters after 2 epochs within the Adam-step bound of the parity tests."""
import os
from types import SimpleNamespace as NS

import numpy as np
import torch

from conftest import GOLDEN

LOGSTD = [-1.4, -3.2]


def _reference_bc(params, loader, eval_loader, episodes, lr):
    from oracle import ref_path as O
    adam = O.AdamState(params, lr, 1e-8, (0.9, 0.999))
    hist = []
    for _ in range(episodes):
        total, nb = 0.0, 0
        for obs, met, act in loader:
            leaf = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
            _, logp, ent, _, _ = O.evaluate_actions(leaf, obs, met, act, True, LOGSTD)
            loss = -logp.mean() - 0 * ent
            grads = torch.autograd.grad(loss, list(leaf.values()), allow_unused=True)
            g = {k: (gr if gr is not None else torch.zeros_like(v)) for (k, v), gr in zip(leaf.items(), grads)}
            adam.step(params, g)
            total += float(loss.detach()); nb += 1
        etotal, ne = 0.0, 0
        with torch.no_grad():
            for obs, met, act in eval_loader:
                _, logp, _, _, _ = O.evaluate_actions(params, obs, met, act, True, LOGSTD)
                etotal += float(-logp.mean()); ne += 1
        hist.append((total / nb, etotal / ne))
    return hist


def test_learn_bc_matches_autograd_restatement(emulated_abi, tmp_path):
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    from gail_carla_b200.expert import DeviceExpertLoader, ExpertDataset
    from gail_carla_b200.learn_bc import learn_bc
    ds = ExpertDataset(os.path.join(GOLDEN, "expert_ds"), routes=[0, 3], n_eps=1)
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    torch.manual_seed(1)
    pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, LOGSTD, False)
    params = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    lr, episodes = 3e-4, 2
    train = DeviceExpertLoader(ds, 4, shuffle=True, drop_last=True, device="cpu")
    val = DeviceExpertLoader(ds, 3, shuffle=False, drop_last=True, device="cpu")
    seen = []
    writer = NS(add_scalar=lambda t, v, s: seen.append((t, s)))
    torch.manual_seed(9)
    hist = learn_bc(pol, "cpu", train, val, episodes=episodes, lr=lr, writer=writer, save_path=str(tmp_path / "bc.pt"))
    # reference-style loop on the same batch sequence (materialised fp32 tuples in the loader's order)
    torch.manual_seed(9)
    class Host:
        def __init__(self, l): self.l = l
        def __iter__(self):
            for b in self.l:
                yield tuple(t.clone() for t in b)
    ref = _reference_bc(params, Host(train), Host(val), episodes, lr)
    np.testing.assert_allclose(np.asarray(hist), np.asarray(ref), rtol=1e-4, atol=1e-6)
    assert seen == [("loss", 0), ("eval_loss", 0), ("loss", 1), ("eval_loss", 1)]
    assert os.path.exists(tmp_path / "bc.pt")
    steps = episodes * len(train)
    for k, v in pol.state_dict().items():
        d = (v.detach().double() - params[k].double()).abs()
        assert d.max().item() <= 2.5 * lr * steps + 1e-3 * params[k].abs().max().item(), k
        assert d.mean().item() <= 0.05 * lr, (k, d.mean().item())


def _bc_golden():
    import json
    return json.load(open(os.path.join(GOLDEN, "learn_bc.json")))


def test_learn_bc_matches_the_unmodified_reference_loop(emulated_abi, tmp_path):
    """The loss / eval_loss series of the UNMODIFIED learn_bc.py:15-72 (tests/golden/make_learn_bc_golden.py: reference Policy,
    reference loop, torch.optim.Adam) vs gail_carla_b200.learn_bc.learn_bc on the CPU statements of the ABI from the same seed
    and batches: same scalar titles / epochs, values within 2e-3 (fp32 re-association through Adam steps of lr 3e-4; the loss
    is -log N(a; mu, sigma = e^-3.2), i.e. (a - mu)^2 amplified ~300x)."""
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    from gail_carla_b200.learn_bc import learn_bc
    gold = _bc_golden(); c = gold["case"]
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    pol = G.Policy(synthetic.OBS_SHAPE, NS(shape=(4,)), NS(shape=(2,)), True, LOGSTD, False)
    train = synthetic.SyntheticExpertLoader(c["n_train"], c["B"], seed=31)
    val = synthetic.SyntheticExpertLoader(c["n_eval"], c["B_eval"], seed=32)
    rows = []
    writer = NS(add_scalar=lambda t, v, s: rows.append([t, float(v), int(s)]))
    learn_bc(pol, "cpu", train, val, episodes=gold["epochs"], writer=writer, save_path=str(tmp_path / "bc.pt"))
    assert [(t, s) for t, _, s in rows] == [(t, s) for t, _, s in gold["scalars"]]
    np.testing.assert_allclose([v for _, v, _ in rows], [v for _, v, _ in gold["scalars"]], rtol=2e-3, atol=1e-4)


def test_reference_bc_pattern_runs_on_the_drop_in_policy(emulated_abi):
    """The reference's own BC step - ``evaluate_actions`` -> ``loss.backward()`` -> ``torch.optim.Adam.step()``
    (learn_bc.py:22,37-45) - written against the drop-in ``Policy``: autograd flows through the engine's hand-derived backward,
    a stock torch optimiser updates the parameters in place and the engine picks the new values up; the loss series equals
    the unmodified reference's."""
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic
    gold = _bc_golden(); c = gold["case"]
    torch.manual_seed(c["seed"]); np.random.seed(c["seed"])
    pol = G.Policy(synthetic.OBS_SHAPE, NS(shape=(4,)), NS(shape=(2,)), True, LOGSTD, False)
    train = synthetic.SyntheticExpertLoader(c["n_train"], c["B"], seed=31)
    val = synthetic.SyntheticExpertLoader(c["n_eval"], c["B_eval"], seed=32)
    optimizer = torch.optim.Adam(pol.parameters(), lr=3e-4)
    series = []
    for epoch in range(gold["epochs"]):
        total, nb = 0.0, 0
        for obs, met, act in train:
            _, logp, entropy, _, _ = pol.evaluate_actions(obs, met, act)
            loss = -logp.mean() - 0 * entropy
            total += float(loss.detach()); nb += 1
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
        etotal, ne = 0.0, 0
        for obs, met, act in val:
            with torch.no_grad():
                _, logp, entropy, _, _ = pol.evaluate_actions(obs, met, act)
            etotal += float(-logp.mean()); ne += 1
        series += [total / nb, etotal / ne]
    np.testing.assert_allclose(series, [v for _, v, _ in gold["scalars"]], rtol=2e-3, atol=1e-4)
