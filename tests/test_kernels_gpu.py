"""GPU parity of the HBM-bound / small kernels: each C-ABI op vs its CPU statement (oracle/abi_emu.py) and, where
the reference produced vectors, vs tests/golden/*.npz (generated from the unmodified reference)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _abi():
    from gail_carla_b200 import _abi
    return _abi


def _emu():
    from oracle import abi_emu
    return abi_emu


def close(a, b, rtol=1e-5, atol=1e-6, what=""):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    if rtol == 0 and atol == 0:
        assert torch.equal(a, b), f"{what}: not bit-identical (max abs diff {(a - b).abs().max().item():.3g})"
        return
    err = ((a - b).abs() / (atol + rtol * b.abs())).max().item() if a.numel() else 0.0
    assert err <= 1.0, f"{what}: max scaled err {err:.3g} (rtol={rtol}, atol={atol})"


@pytest.mark.parametrize("name", ["gae_64x4", "gae_2048x16", "gae_33x1"])
def test_gae_matches_reference_golden(name):
    """compute_returns + advantage normalisation vs the reference's own outputs (tolerance: fp32 re-association of the
    scan, rtol 1e-5 / atol 2e-6 on returns; 1e-4 on normalised advantages)."""
    A = _abi()
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    r, v, m = (torch.from_numpy(z[k]).to(DEV) for k in ("gail_rewards", "value_preds", "masks"))
    T, N = r.shape[:2]
    ret = torch.zeros_like(v); adv = torch.zeros_like(r); stats = torch.zeros(4, dtype=torch.float64, device=DEV)
    A.gae_returns(r, v, m, ret, float(z["gamma"]), float(z["gae_lambda"]), adv, stats)
    close(ret, torch.from_numpy(z["returns"]), 1e-5, 2e-6, "returns")
    assert float(ret[-1].abs().max()) == 0.0
    out = torch.zeros_like(r)
    A.adv_normalize(ret, v, stats, out, T * N)
    close(out, torch.from_numpy(z["adv_norm"]), 1e-4, 1e-5, "adv_norm")
    stats2 = torch.zeros(4, dtype=torch.float64, device=DEV)
    A.adv_stats(ret, v, stats2, T * N)
    close(stats2[:3], stats[:3], 1e-6, 1e-6, "stats")


@pytest.mark.parametrize("T,N", [(1, 1), (7, 3), (128, 1), (513, 40), (1024, 64), (512, 256), (37, 1000)])
def test_gae_shapes_vs_sequential(T, N):
    A, E = _abi(), _emu()
    g = torch.Generator().manual_seed(T * 1000 + N)
    r = torch.nn.functional.softplus(torch.randn(T, N, 1, generator=g)); v = torch.randn(T + 1, N, 1, generator=g)
    m = (torch.rand(T + 1, N, 1, generator=g) > 0.05).float()
    ret_ref = torch.zeros_like(v); E.gae_returns(r, v, m, ret_ref, 0.99, 0.95)
    ret = torch.zeros(T + 1, N, 1, device=DEV)
    A.gae_returns(r.to(DEV), v.to(DEV), m.to(DEV), ret, 0.99, 0.95)
    close(ret, ret_ref, 1e-5, 3e-6, f"gae {T}x{N}")


def test_gae_constant_reward_closed_form():
    """no terminations, r=1, V=0: gae_t = sum_k (gamma*lam)^k -> geometric series (size-independent property)."""
    A = _abi()
    T, N = 4096, 512
    r = torch.ones(T, N, 1, device=DEV); v = torch.zeros(T + 1, N, 1, device=DEV); m = torch.ones(T + 1, N, 1, device=DEV)
    ret = torch.zeros_like(v)
    A.gae_returns(r, v, m, ret, 0.99, 0.95)
    gl = 0.99 * 0.95
    k = torch.arange(T, 0, -1, dtype=torch.float64)
    expect = (1 - gl ** k) / (1 - gl)
    close(ret[:T, 7, 0], expect, 1e-5, 1e-6, "closed form")


@pytest.mark.parametrize("B", [1000, 1003, 3])     # 4 samples per thread + ragged tail
@pytest.mark.parametrize("mode,use_adv", [(0, True), (0, False), (1, False), (2, False)])
def test_ppo_loss_vs_autograd(mode, use_adv, B):
    A, E = _abi(), _emu()
    g = torch.Generator().manual_seed(5 + mode)
    head = torch.randn(B, 4, generator=g); act = torch.randn(B, 2, generator=g) * 0.3
    act[:, 1] = act[:, 1].abs()
    logstd = (-1.4, -3.2)
    # old log-probs near the new ones so ratios straddle the clip range
    _, _, _, lp = E._head_tail(head, act, logstd, True)
    olp = lp + torch.randn(B, generator=g) * 0.08
    vo = head[:, 0] + torch.randn(B, generator=g) * 0.15; ret = torch.randn(B, generator=g)
    adv = torch.randn(B, generator=g)
    stats = torch.tensor([3.0, 1200.0, float(B), 0.0], dtype=torch.float64)
    outs = {}
    for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
        t = lambda x: x.clone().to(dev)
        dh = torch.zeros(B, 4, device=dev); ov = torch.zeros(B, device=dev); ol = torch.zeros(B, device=dev)
        acc = torch.zeros(4, dtype=torch.float64, device=dev)
        mod.ppo_loss(t(head), t(act), t(olp), t(vo), t(ret), t(adv) if use_adv else None, t(stats), dh, ov, ol, acc, B, logstd,
                     True, 0.1, 0.5, 0.7, mode)
        outs[tag] = (dh, ov, ol, acc)
    close(outs["gpu"][1], outs["ref"][1], 1e-6, 1e-6, "value")
    close(outs["gpu"][2], outs["ref"][2], 2e-5, 2e-5, "logp")
    if mode != 2:
        close(outs["gpu"][0], outs["ref"][0], 1e-4, 1e-7, "d_head")
        close(outs["gpu"][3][:3], outs["ref"][3][:3], 1e-5, 1e-4, "loss sums")


def test_policy_act_and_reward_and_welford():
    A, E = _abi(), _emu()
    B = 333
    g = torch.Generator().manual_seed(9)
    head = torch.randn(B, 4, generator=g); noise = torch.randn(B, 2, generator=g)
    for nz in (None, noise):
        res = {}
        for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
            v = torch.zeros(B, device=dev); a = torch.zeros(B, 2, device=dev); lp = torch.zeros(B, device=dev)
            mod.policy_act(head.to(dev), None if nz is None else nz.to(dev), v, a, lp, B, (-1.4, -3.2), True)
            res[tag] = (v, a, lp)
        for x, y in zip(res["gpu"], res["ref"]):
            close(x, y, 1e-5, 1e-5, "act")
    d = torch.randn(1000, generator=g) * 4
    r_ref = torch.zeros(1000); E.reward_epilogue(d, r_ref, 1000)
    r = torch.zeros(1000, device=DEV); A.reward_epilogue(d.to(DEV), r, 1000)
    close(r, r_ref, 2e-5, 1e-6, "reward")
    # RunningMeanStd golden (common/running_mean_std.py) - float64 state
    z = np.load(os.path.join(GOLDEN, "rms.npz"))
    st = torch.tensor([0.0, 1.0, 1e-4], dtype=torch.float64, device=DEV); scratch = torch.zeros(2, dtype=torch.float64, device=DEV)
    for i in range(3):
        A.welford_merge(st, torch.from_numpy(z[f"x{i}"]).float().to(DEV), scratch)
        ref = torch.from_numpy(z["hist"][i])
        close(st, ref, 1e-5, 1e-6, f"rms step {i}")   # inputs are cast to fp32 on the device side


def test_gather_mixup_metrics():
    A, E = _abi(), _emu()
    g = torch.Generator().manual_seed(11)
    rows, B = 9, 5
    src = torch.rand(rows, 3, 192, 192, generator=g)
    idx = torch.tensor([8, 0, 3, 3, 7])
    for ix in (idx, None):
        o_ref = torch.zeros(B, 96, 96, 16); E.gather_obs_s2d(src, ix, o_ref, B)
        o = torch.zeros(B, 96, 96, 16, device=DEV); A.gather_obs_s2d(src.to(DEV), None if ix is None else ix.to(DEV), o, B)
        close(o, o_ref, 1e-6, 1e-6, "gather_obs")
    # uint8 table of a device-resident expert data set: decoding first (uint8/255 -> fp32 rows) and gathering those rows
    # with the fp32 kernel must give the same bits as the fused uint8 kernel; both within rounding of the CPU statement
    src8 = torch.randint(0, 256, (rows, 3, 192, 192), generator=g, dtype=torch.uint8)
    for ix in (idx, None):
        ixd = None if ix is None else ix.to(DEV)
        o8 = torch.zeros(B, 96, 96, 16, device=DEV); A.gather_obs_s2d(src8.to(DEV), ixd, o8, B)
        of = torch.zeros(B, 96, 96, 16, device=DEV); A.gather_obs_s2d((src8.float() / 255).to(DEV), ixd, of, B)
        close(o8, of, 0, 0, "gather_obs uint8 vs decoded fp32")
        o_ref = torch.zeros(B, 96, 96, 16); E.gather_obs_s2d(src8, ix, o_ref, B)
        close(o8, o_ref, 1e-6, 1e-6, "gather_obs uint8")
        assert (o8.view(B, 96, 96, 4, 4)[..., 3] == 1).all(), "pad channel must hold 1.0"
    # fused critic minibatch: expert rows | policy rows | mix-up in one pass == the two uint8 gathers + gc_mixup, bit for bit
    srcp = torch.randint(0, 256, (rows, 3, 192, 192), generator=g, dtype=torch.uint8)
    idx_p = torch.tensor([1, 1, 6, 2, 0]); al8 = torch.rand(B, generator=g)
    per = 96 * 96 * 16
    fused = torch.zeros(3 * B, per, device=DEV)
    A.gather_pair_mix(src8.to(DEV), idx.to(DEV), srcp.to(DEV), idx_p.to(DEV), al8.to(DEV), fused, B)
    sep = torch.zeros(3 * B, per, device=DEV)
    A.gather_obs_s2d(src8.to(DEV), idx.to(DEV), sep, B); A.gather_obs_s2d(srcp.to(DEV), idx_p.to(DEV), sep[B:], B)
    A.mixup(sep, sep[B:], al8.to(DEV), sep[2 * B:], B, per)
    close(fused, sep, 0, 0, "gather_pair_mix vs gather + gather + mixup")
    fused_none = torch.zeros(3 * B, per, device=DEV)
    A.gather_pair_mix(src8.to(DEV), None, srcp.to(DEV), None, al8.to(DEV), fused_none, B)
    A.gather_obs_s2d(src8.to(DEV), None, sep, B); A.gather_obs_s2d(srcp.to(DEV), None, sep[B:], B)
    A.mixup(sep, sep[B:], al8.to(DEV), sep[2 * B:], B, per)
    close(fused_none, sep, 0, 0, "gather_pair_mix (no index) vs separate kernels")
    s2 = torch.randn(rows, 4, generator=g)
    o_ref = torch.zeros(B, 8); E.gather_rows(s2, idx, o_ref, B, 4, 8)
    o = torch.zeros(B, 8, device=DEV); A.gather_rows(s2.to(DEV), idx.to(DEV), o, B, 4, 8)
    close(o, o_ref, 0, 0, "gather_rows")
    xe = torch.randn(B, 4096, generator=g); xp = torch.randn(B, 4096, generator=g); al = torch.rand(B, generator=g)
    m_ref = torch.zeros(B, 4096); E.mixup(xe, xp, al, m_ref, B, 4096)
    m = torch.zeros(B, 4096, device=DEV); A.mixup(xe.to(DEV), xp.to(DEV), al.to(DEV), m, B, 4096)
    close(m, m_ref, 1e-6, 1e-7, "mixup")
    met = torch.cat([torch.randn(B, 2, generator=g) * 1e-3, torch.rand(B, 1, generator=g) * 8,
                     torch.randint(1, 7, (B, 1), generator=g).float()], 1)
    met2 = torch.cat([torch.randn(B, 2, generator=g) * 1e-3, torch.rand(B, 1, generator=g) * 8,
                      torch.randint(1, 7, (B, 1), generator=g).float()], 1)
    act = torch.randn(B, 2, generator=g); act2 = torch.randn(B, 2, generator=g); emb = torch.randn(10, 8, generator=g)
    for mix in (False, True):
        kw = dict(action=act, metrics2=met2 if mix else None, action2=act2 if mix else None, alpha=al if mix else None)
        f_ref = torch.zeros(B, 40); E.metrics_features(met, emb, f_ref[:, 8:], 40, 32, B, **kw)
        f = torch.zeros(B, 40, device=DEV)
        A.metrics_features(met.to(DEV), emb.to(DEV), f[:, 8:], 40, 32, B, **{k: (None if v is None else v.to(DEV)) for k, v in kw.items()})
        close(f, f_ref, 2e-6, 1e-6, "metrics_features")
        df = torch.randn(B, 40, generator=g)
        de_ref = torch.zeros(10, 8); E.metrics_features_bwd(met, df[:, 8:], 40, de_ref, B, met2 if mix else None, al if mix else None)
        de = torch.zeros(10, 8, device=DEV)
        A.metrics_features_bwd(met.to(DEV), df.to(DEV)[:, 8:], 40, de, B, met2.to(DEV) if mix else None, al.to(DEV) if mix else None)
        close(de, de_ref, 1e-5, 1e-6, "metrics_features_bwd")


def test_small_linear_seed_penalty_colsum_reduce():
    A, E = _abi(), _emu()
    g = torch.Generator().manual_seed(13)
    B, K = 300, 256
    for N, ldy in ((3, 4), (1, 1)):
        x = torch.randn(B, K + 4, generator=g); w = torch.randn(N, K, generator=g); b = torch.randn(N, generator=g)
        dy = torch.randn(B, ldy, generator=g)
        res = {}
        for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
            y = torch.zeros(B, ldy, device=dev); dx = torch.zeros(B, K + 4, device=dev)
            dw = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
            mod.small_linear_fwd(x.to(dev), K + 4, w.to(dev), b.to(dev), y, ldy, B, N, K)
            mod.small_linear_bwd(x.to(dev), K + 4, w.to(dev), dy.to(dev), ldy, dx, K + 4, dw, db, B, 200, N, K, 0.2)
            res[tag] = (y, dx, dw, db)
        for a_, b_, nm in zip(res["gpu"], res["ref"], ("y", "dx", "dw", "db")):
            close(a_, b_, 1e-4, 1e-4, "small_linear " + nm)
    d = torch.randn(3 * B, generator=g)
    res = {}
    for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
        dd = torch.zeros(3 * B, device=dev); acc = torch.zeros(4, dtype=torch.float64, device=dev)
        mod.disc_loss_seed(d.to(dev), dd, acc, B)
        res[tag] = (dd, acc)
    close(res["gpu"][0], res["ref"][0], 1e-5, 1e-8, "seed"); close(res["gpu"][1], res["ref"][1], 1e-6, 1e-5, "seed acc")
    per = 96 * 96 * 16
    gg = torch.randn(4, per, generator=g) * 0.01
    sc = tuple(1.0 / s for s in (0.229, 0.224, 0.225))
    res = {}
    for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
        u = torch.zeros(4, per, device=dev); acc = torch.zeros(2, dtype=torch.float64, device=dev)
        mod.grad_penalty(gg.to(dev), u, acc, 4, per, 10.0, sc)
        res[tag] = (u, acc)
    close(res["gpu"][0], res["ref"][0], 1e-4, 1e-7, "gp u"); close(res["gpu"][1][:1], res["ref"][1][:1], 1e-5, 1e-6, "gp acc")
    x = torch.randn(1000, 72, generator=g)
    o_ref = torch.ones(64); E.colsum(x, 72, 1000, 64, o_ref)
    o = torch.ones(64, device=DEV); A.colsum(x.to(DEV), 72, 1000, 64, o)
    close(o, o_ref, 1e-4, 1e-4, "colsum")
    part = torch.randn(3, 50, 24, generator=g); bias = torch.randn(20, generator=g); ms = torch.randn(50, 24, generator=g)
    for epi in (0, 1, 2, 3):
        o_ref = torch.zeros(50, 24); E.splitk_reduce(part, 3, 50, 20, 24, bias, ms, 24, o_ref, 24, epi, 0.2)
        o = torch.zeros(50, 24, device=DEV)
        A.splitk_reduce(part.to(DEV), 3, 50, 20, 24, bias.to(DEV), ms.to(DEV), 24, o, 24, epi, 0.2)
        close(o, o_ref, 1e-5, 1e-6, f"splitk epi{epi}")


def test_weight_layouts_roundtrip_and_adam():
    A, E = _abi(), _emu()
    g = torch.Generator().manual_seed(17)
    for Cout, Cin, l1 in ((32, 3, 1), (64, 32, 0), (256, 128, 0)):
        w = torch.randn(Cout, Cin, 4, 4, generator=g)
        n = Cout * (16 if l1 else Cin) * (4 if l1 else 16)
        n = 2048 if l1 else Cout * Cin * 16
        res = {}
        for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
            wf = torch.zeros(n, device=dev); wd = torch.zeros(n, device=dev)
            mod.prep_conv_weight(w.to(dev), wf, wd, Cout, Cin, l1)
            part = torch.stack([wf, 2 * wf]); dw = torch.zeros(Cout, Cin, 4, 4, device=dev)
            if l1:   # the pad-channel column of tap (0,0) carries the bias gradient (pad channel of X0 = 1.0)
                part = part.clone(); part.view(2, 32, 64)[:, :, 3] = torch.arange(64, dtype=torch.float32, device=dev).view(2, 32)
            db = torch.zeros(Cout, device=dev)
            mod.unprep_conv_wgrad(part, 2, dw, Cout, Cin, l1, db if l1 else None)
            if l1:
                close(db.cpu(), (torch.arange(32.) + torch.arange(32., 64.)), 0, 0, "conv1 bias-grad column")
            res[tag] = (wf, wd, dw)
        for a_, b_ in zip(res["gpu"], res["ref"]):
            close(a_, b_, 0, 0, "conv layouts")
        close(res["gpu"][2], 3 * w, 1e-6, 1e-6, "conv layout roundtrip")
    out, tail, ld = 8, 13, 25632
    w = torch.randn(out, 25600 + tail, generator=g)
    res = {}
    for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
        wg = torch.full((out, ld), 5.0, device=dev); mod.prep_fc1_weight(w.to(dev), wg, out, tail, ld)
        dw = torch.zeros(out, 25600 + tail, device=dev); mod.unprep_fc1_wgrad(torch.stack([wg, wg]), 2, dw, out, tail, ld)
        res[tag] = (wg, dw)
    close(res["gpu"][0], res["ref"][0], 0, 0, "fc1 prep"); close(res["gpu"][1], 2 * w, 1e-6, 1e-6, "fc1 roundtrip")
    n = 100003
    p = torch.randn(n, generator=g); gr = torch.randn(n, generator=g) * 0.01
    res = {}
    for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
        P, G = p.clone().to(dev), gr.clone().to(dev); M = torch.zeros(n, device=dev); V = torch.zeros(n, device=dev)
        for t in (1, 2, 3):
            ss = torch.zeros(1, dtype=torch.float64, device=dev)
            mod.grad_sumsq(G, n, ss)
            mod.clip_adam(P, G, M, V, n, ss, 0.5, 1e-3, 0.9, 0.99, 1e-8, 1 - 0.9 ** t, 1 - 0.99 ** t)
        res[tag] = (P, M, V, ss)
    for a_, b_, nm in zip(res["gpu"], res["ref"], "PMVs"):
        close(a_, b_, 1e-5, 1e-7, "adam " + nm)
    # round-2 arguments: gradient scale (1/world after a NCCL sum), zero_grad folded into the Adam pass, hyper-parameters
    # read from device memory (CUDA-graph replay)
    res = {}
    for tag, mod, dev in (("ref", E, "cpu"), ("gpu", A, DEV)):
        P, G = p.clone().to(dev), (gr * 4).clone().to(dev); M = torch.zeros(n, device=dev); V = torch.zeros(n, device=dev)
        ss = torch.zeros(1, dtype=torch.float64, device=dev)
        mod.grad_sumsq(G, n, ss, 0.25)
        hyper = torch.tensor([1e-3, 1 - 0.9, 1 - 0.99], device=dev)
        mod.clip_adam(P, G, M, V, n, ss, 0.5, 123.0, 0.9, 0.99, 1e-8, 0.5, 0.5, 0.25, True, hyper)
        res[tag] = (P, M, V, ss, G)
    for a_, b_, nm in zip(res["gpu"], res["ref"], "PMVsG"):
        close(a_, b_, 1e-5, 1e-7, "adam (scaled, zeroing, device hyper) " + nm)
    assert float(res["gpu"][4].abs().max()) == 0.0
    ss1 = torch.zeros(1, dtype=torch.float64, device=DEV); A.grad_sumsq(gr.to(DEV), n, ss1)
    close(res["gpu"][3], ss1, 1e-6, 0, "scaled sumsq == sumsq of the scaled gradient")
