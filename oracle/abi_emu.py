"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU statement (torch fp32) of what each C-ABI entry point in include/gail_carla_b200.h computes, with the same
call signatures as the wrappers in ``gail_carla_b200/_abi.py`` but operating on CPU tensors.  Used
  * by the GPU tests: each CUDA op is compared with its statement here on the same inputs;
  * by the CPU tests: ``tests/conftest.py::emulated_abi`` swaps these functions in for the ctypes wrappers so the
    host-side classes (layout bookkeeping, hand-derived backward passes, optimiser plumbing) can be checked against
    ``oracle.ref_path`` / the golden vectors in a container without a GPU.
Pointers + pitches of the ABI are modelled as views on the tensor's storage starting at its storage offset.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle.ref_path import NORM_MEAN, NORM_STD

EPI_STORE, EPI_BIAS_LRELU, EPI_BIAS, EPI_MASK = 0, 1, 2, 3


def _v(t, shape, strides):
    """View `shape`/`strides` (in elements) on t's storage, starting at t's first element."""
    return torch.as_strided(t, tuple(int(s) for s in shape), tuple(int(s) for s in strides), t.storage_offset())


def _v2(t, rows, cols, ld):
    return _v(t, (rows, cols), (ld, 1))


def _slope_mask(src, slope):
    return torch.where(src > 0, torch.ones_like(src), torch.full_like(src, slope))


# ------------------------------------------------------------------ rollout maths
def gae_returns(gail_rewards, value_preds, masks, returns, gamma, gae_lambda, adv_raw=None, stats=None):
    T, N = gail_rewards.shape[0], gail_rewards.shape[1]
    r, v, m, out = (x.view(-1, N) for x in (gail_rewards, value_preds, masks, returns))
    gae = torch.zeros(N)
    for t in reversed(range(T)):
        delta = r[t] + gamma * v[t + 1] * m[t + 1] - v[t]
        gae = delta + gamma * gae_lambda * m[t + 1] * gae
        out[t] = gae + v[t]
    a = out[:T] - v[:T]
    if adv_raw is not None:
        adv_raw.view(-1, N)[:T] = a
    if stats is not None:
        stats[0] = a.double().sum(); stats[1] = (a.double() ** 2).sum(); stats[2] = T * N


def adv_stats(returns, value_preds, stats, n):
    a = (returns.reshape(-1)[:n] - value_preds.reshape(-1)[:n]).double()
    stats[0] = a.sum(); stats[1] = (a ** 2).sum(); stats[2] = n


def _moments(stats):
    s, q, n = (float(x) for x in stats[:3])
    mean = s / n
    var = max((q - s * mean) / (n - 1), 0.0)
    return mean, 1.0 / (math.sqrt(var) + 1e-5)


def adv_normalize(returns, value_preds, stats, out, n):
    mean, inv = _moments(stats)
    out.view(-1)[:n] = ((returns.reshape(-1)[:n] - value_preds.reshape(-1)[:n]) - mean) * inv


def _head_tail(head, actions, logstd, activation):
    v = head[:, 0]
    mu0 = torch.tanh(head[:, 1]) if activation else head[:, 1]
    mu1 = torch.sigmoid(head[:, 2]) if activation else head[:, 2]
    ls = torch.tensor([float(logstd[0]), float(logstd[1])])
    var = torch.exp(ls) ** 2
    c = 0.5 * math.log(2 * math.pi)
    logp = (-(actions[:, 0] - mu0) ** 2 / (2 * var[0]) - ls[0] - c) + (-(actions[:, 1] - mu1) ** 2 / (2 * var[1]) - ls[1] - c)
    return v, mu0, mu1, logp


@torch.enable_grad()
def ppo_loss(head_out, actions, old_logp, value_old, returns, adv, adv_stats_, d_head, out_value, out_logp, loss_acc, B,
             logstd, activation, clip, value_coef, action_weight, mode, clipped_value=True, norm=None):
    h = head_out.view(-1, 4)[:B].detach().clone().requires_grad_(mode != 2)
    a = actions.view(-1, 2)[:B]
    v, _, _, logp = _head_tail(h, a, logstd, activation)
    if out_value is not None:
        out_value.view(-1)[:B] = v.detach()
    if out_logp is not None:
        out_logp.view(-1)[:B] = logp.detach()
    if mode == 2:
        return
    if mode == 0:
        R, vo, olp = returns.view(-1)[:B], value_old.view(-1)[:B], old_logp.view(-1)[:B]
        if adv is not None:
            A = adv.view(-1)[:B]
        else:
            mean, inv = _moments(adv_stats_)
            A = ((R - vo) - mean) * inv
        ratio = torch.exp(logp - olp)
        neg_min = -torch.min(ratio * A, torch.clamp(ratio, 1 - clip, 1 + clip) * A)
        if clipped_value:
            vc = vo + (v - vo).clamp(-clip, clip)
            vl = 0.5 * torch.max((v - R) ** 2, (vc - R) ** 2)
        else:
            vl = 0.5 * (R - v) ** 2
        nb = B if norm is None else norm
        (value_coef * vl.sum() / nb + action_weight * neg_min.sum() / nb).backward()
        if loss_acc is not None:
            loss_acc[0] += vl.detach().double().sum(); loss_acc[1] += neg_min.detach().double().sum()
    else:
        (action_weight * (-logp).sum() / (B if norm is None else norm)).backward()
        if loss_acc is not None:
            loss_acc[2] += (-logp).detach().double().sum()
    g = h.grad.clone()
    g[:, 3] = 0
    d_head.view(-1, 4)[:B] = g


def policy_act(head_out, noise, value, action, logp, B, logstd, activation):
    h = head_out.view(-1, 4)[:B]
    v, mu0, mu1, _ = _head_tail(h, torch.zeros(B, 2), logstd, activation)
    a = torch.stack([mu0, mu1], 1)
    if noise is not None:
        a = a + torch.exp(torch.tensor([float(logstd[0]), float(logstd[1])])) * noise.view(-1, 2)[:B]
    _, _, _, lp = _head_tail(h, a, logstd, activation)
    value.view(-1)[:B] = v; action.view(-1, 2)[:B] = a; logp.view(-1)[:B] = lp


def welford_merge(state, x, scratch2):
    xs = x.double().reshape(-1)
    n = xs.numel()
    bm, bv = xs.mean(), xs.var(unbiased=False)
    mean, var, count = state[0].clone(), state[1].clone(), state[2].clone()
    delta, tot = bm - mean, count + n
    state[0] = mean + delta * n / tot
    state[1] = (var * count + bv * n + delta ** 2 * count * n / tot) / tot
    state[2] = tot


# ------------------------------------------------------------------ data movement / small stages
def gather_obs_s2d(src, idx, out, B):
    rows = src.reshape(-1, 3, 192, 192)
    x = rows[idx[:B]] if idx is not None else rows[:B]
    if x.dtype == torch.uint8:          # device-resident expert table: ToTensor() = uint8 / 255 (algo/wdgail.py:222-227)
        x = x.float() / 255.0
    mean = torch.tensor(NORM_MEAN).view(1, 3, 1, 1); std = torch.tensor(NORM_STD).view(1, 3, 1, 1)
    x = (x - mean) / std
    x4 = torch.cat([x, torch.ones(B, 1, 192, 192)], 1)                       # [B,4,192,192], pad channel = 1
    x4 = x4.view(B, 4, 96, 2, 96, 2).permute(0, 2, 4, 3, 5, 1)               # b,Y,X,dy,dx,c
    out.view(-1)[:B * 96 * 96 * 16] = x4.reshape(-1)


def gather_pair_mix(src_e, idx_e, src_p, idx_p, alpha, out, B):
    per = 96 * 96 * 16
    o = out.view(-1)
    gather_obs_s2d(src_e, idx_e, o[:B * per], B)
    gather_obs_s2d(src_p, idx_p, o[B * per:2 * B * per], B)
    mixup(o[:B * per], o[B * per:2 * B * per], alpha, o[2 * B * per:3 * B * per], B, per)


def gather_rows(src, idx, out, B, width, ldo):
    rows = src.reshape(-1, width)
    _v2(out, B, width, ldo).copy_(rows[idx[:B]] if idx is not None else rows[:B])


def mixup(xe, xp, alpha, out, B, per_sample):
    a = alpha.view(-1)[:B].view(B, 1)
    out.view(-1)[:B * per_sample] = (a * xe.reshape(-1)[:B * per_sample].view(B, -1)
                                     + (1 - a) * xp.reshape(-1)[:B * per_sample].view(B, -1)).reshape(-1)


def _mixed(m, m2, alpha, B, w):
    m = m.reshape(-1, w)[:B]
    if m2 is not None:
        a = alpha.view(-1)[:B].view(B, 1)
        m = a * m + (1 - a) * m2.reshape(-1, w)[:B]
    return m


def metrics_features(metrics, emb, out, ldo, pad, B, action=None, metrics2=None, action2=None, alpha=None):
    m = _mixed(metrics, metrics2, alpha, B, 4)
    x, y = m[:, 0], m[:, 1]
    cols = [1000 * x, 1000 * y, 1000 * torch.sqrt(x * x + y * y), 0.3 * torch.atan2(y, x), 0.1 * m[:, 2]]
    idx = m[:, 3].long().clamp(0, 9)
    f = torch.cat([torch.stack(cols, 1), emb.view(10, 8)[idx]], 1)
    if action is not None:
        f = torch.cat([f, _mixed(action, action2, alpha, B, 2)], 1)
    o = _v2(out, B, pad, ldo)
    o.zero_()
    o[:, :f.shape[1]] = f.detach()


def metrics_features_bwd(metrics, d_feat, ldf, d_emb, B, metrics2=None, alpha=None):
    m = _mixed(metrics, metrics2, alpha, B, 4)
    idx = m[:, 3].long().clamp(0, 9)
    d_emb.view(10, 8).index_add_(0, idx, _v2(d_feat, B, 13, ldf)[:, 5:13])


def small_linear_fwd(x, ldx, w, bias, y, ldy, B, N, K):
    r = _v2(x, B, K, ldx) @ w.view(N, K).t()
    if bias is not None:
        r = r + bias.view(-1)[:N]
    _v2(y, B, N, ldy).copy_(r)


def small_linear_bwd(x, ldx, w, dy, lddy, dx, lddx, dw, db, B, B_params, N, K, slope):
    X, DY = _v2(x, B, K, ldx), _v2(dy, B, N, lddy)
    if dx is not None:
        d = DY @ w.view(N, K)
        if slope >= 0:
            d = d * _slope_mask(X, slope)
        _v2(dx, B, K, lddx).copy_(d)
    if dw is not None and B_params > 0:
        dw.view(N, K).add_(DY[:B_params].t() @ X[:B_params])
        if db is not None:
            db.view(-1)[:N].add_(DY[:B_params].sum(0))


def disc_loss_seed(d, dd, acc, B, norm=None):
    d = d.view(-1)
    te, tp = torch.tanh(d[:B]), torch.tanh(d[B:2 * B])
    o = dd.view(-1)
    nb = B if norm is None else norm
    o[:B] = -(1 - te * te) / nb
    o[B:2 * B] = (1 - tp * tp) / nb
    o[2 * B:3 * B] = 1.0
    acc[0] += d[:B].double().sum(); acc[1] += d[B:2 * B].double().sum()
    acc[2] += te.double().sum(); acc[3] += tp.double().sum()


def grad_penalty(g, u, acc, B, per_sample, lambda_, scales, norm=None):
    G = g.reshape(-1)[:B * per_sample].view(B, -1, 4)
    s = torch.tensor([scales[0], scales[1], scales[2], 0.0])
    nrm = torch.sqrt(((G * s).double() ** 2).sum((1, 2))).float()
    acc[0] += ((nrm - 1).double() ** 2).sum()
    coef = torch.where(nrm > 0, lambda_ * 2 * (nrm - 1) / ((B if norm is None else norm) * nrm), torch.zeros_like(nrm))
    u.view(-1)[:B * per_sample] = (coef.view(B, 1, 1) * (s * s) * G).reshape(-1)


def reward_epilogue(d, reward, n):
    reward.view(-1)[:n] = -(1 - torch.sigmoid(d.reshape(-1)[:n])).log()


def colsum(x, ld, rows, Cc, out):
    out.view(-1)[:Cc].add_(_v2(x, rows, Cc, ld).sum(0))


def splitk_reduce(part, splits, M, N, ldp, bias, mask_src, ldm, out, ldo, epilogue, slope):
    s = _v(part, (splits, M, N), (M * ldp, ldp, 1)).sum(0)
    if epilogue in (EPI_BIAS_LRELU, EPI_BIAS):
        s = s + bias.view(-1)[:N]
    if epilogue == EPI_BIAS_LRELU:
        s = F.leaky_relu(s, slope)
    if epilogue == EPI_MASK:
        s = s * _slope_mask(_v2(mask_src, M, N, ldm), slope)
    _v2(out, M, N, ldo).copy_(s)


# ------------------------------------------------------------------ parameter layouts / optimiser
def _conv1_fprop_layout(w):
    """[32,3,4,4] -> [32][ky2][px][dy][dx][c4]"""
    w4 = torch.cat([w, torch.zeros(32, 1, 4, 4)], 1)                          # n,c,ky,kx
    return w4.view(32, 4, 2, 2, 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(32, 64)  # n,ky2,px,dy,dx,c


def prep_conv_weight(w, w_fprop, w_dgrad, Cout, Cin, layer1):
    w = w.detach().view(Cout, Cin, 4, 4)
    if layer1:
        wf = _conv1_fprop_layout(w)
        w_fprop.view(-1)[:wf.numel()] = wf.reshape(-1)
        if w_dgrad is not None:   # [q][a][b'][n]
            w_dgrad.view(-1)[:wf.numel()] = wf.view(32, 2, 2, 16).permute(3, 1, 2, 0).reshape(-1)
    else:
        w_fprop.view(-1)[:w.numel()] = w.permute(0, 2, 3, 1).reshape(-1)
        if w_dgrad is not None:   # [py,px][c][a][b'][n] with ky = py + 2a, kx = px + 2b'
            t = w.view(Cout, Cin, 2, 2, 2, 2)                                   # n,c,a,py,b',px
            w_dgrad.view(-1)[:w.numel()] = t.permute(3, 5, 1, 2, 4, 0).reshape(-1)


def unprep_conv_wgrad(part, splits, dw, Cout, Cin, layer1, dbias=None):
    if layer1:
        s = part.reshape(-1)[:splits * 2048].view(splits, 32, 2, 2, 2, 2, 4).sum(0)   # n,ky2,px,dy,dx,c
        dw.view(32, 3, 4, 4).copy_(s.permute(0, 5, 1, 3, 2, 4).reshape(32, 4, 4, 4)[:, :3])
        if dbias is not None:   # pad channel of tap (0,0): sum over pixels of dy[n] * 1.0
            dbias.view(-1)[:32].copy_(s[:, 0, 0, 0, 0, 3])
    else:
        n = Cout * Cin * 16
        s = part.reshape(-1)[:splits * n].view(splits, Cout, 4, 4, Cin).sum(0)
        dw.view(Cout, Cin, 4, 4).copy_(s.permute(0, 3, 1, 2))


def prep_fc1_weight(w, w_gemm, out, tail, ld):
    w = w.detach().view(out, 25600 + tail)
    g = _v2(w_gemm, out, ld, ld)
    g.zero_()
    g[:, :25600] = w[:, :25600].view(out, 256, 100).permute(0, 2, 1).reshape(out, 25600)
    g[:, 25600:25600 + tail] = w[:, 25600:]


def unprep_fc1_wgrad(part, splits, dw, out, tail, ld):
    s = _v(part, (splits, out, ld), (out * ld, ld, 1)).sum(0)
    d = dw.view(out, 25600 + tail)
    d[:, :25600] = s[:, :25600].view(out, 100, 256).permute(0, 2, 1).reshape(out, 25600)
    d[:, 25600:] = s[:, 25600:25600 + tail]


def zero_block(t, rows=1, width=None, pitch=None):
    width = t.numel() if width is None else width
    pitch = width if pitch is None else pitch
    _v(t, (rows, width), (pitch, 1)).zero_()


def grad_sumsq(grad, n, sumsq, grad_scale=1.0):
    sumsq[0] = ((grad.reshape(-1)[:n].double() * grad_scale) ** 2).sum()


def clip_adam(param, grad, exp_avg, exp_avg_sq, n, sumsq, max_norm, lr, beta1, beta2, eps, bc1, bc2, grad_scale=1.0,
              zero_grad=False, dev_hyper=None):
    if dev_hyper is not None:
        lr, bc1, bc2 = (float(x) for x in dev_hyper[:3])
    p, g0, m, v = (t.view(-1)[:n] for t in (param, grad, exp_avg, exp_avg_sq))
    g = g0 * grad_scale
    if max_norm is not None and max_norm >= 0:
        total = torch.sqrt(sumsq[0]).float()
        g = g * torch.clamp(max_norm / (total + 1e-6), max=1.0)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    p.addcdiv_(m, v.sqrt() / math.sqrt(bc2) + eps, value=-lr / bc1)
    if zero_grad:
        g0.zero_()


# ------------------------------------------------------------------ dense contractions
def _in_view(g, x):
    return _v(x, (g.B, g.Hp, g.Wp, g.Cin), (g.in_batch_stride, g.Wp * g.Cin, g.Cin, 1))


def _out_view(g, y):
    return _v(y, (g.B, g.OHp, g.OWp, g.Cout), (g.out_batch_stride, g.OWp * g.Cout, g.Cout, 1))


def conv_fprop(geom, x, w, bias, y, epilogue, slope=0.2, mask_src=None, mask_bits=None):   # mask_bits: accelerator only, ignored
    g = geom
    xin = _in_view(g, x)[:, :g.H, :g.W].permute(0, 3, 1, 2)
    wt = w.reshape(-1)[:g.Cout * g.KH * g.KW * g.Cin].view(g.Cout, g.KH, g.KW, g.Cin).permute(0, 3, 1, 2)
    r = F.conv2d(xin, wt, None, stride=g.S)[:, :, :g.OH, :g.OW].permute(0, 2, 3, 1)
    if epilogue in (EPI_BIAS_LRELU, EPI_BIAS):
        r = r + bias.view(-1)[:g.Cout]
    if epilogue == EPI_BIAS_LRELU:
        r = F.leaky_relu(r, slope)
    if epilogue == EPI_MASK:
        r = r * _slope_mask(_out_view(g, mask_src)[:, :g.OH, :g.OW], slope)
    _out_view(g, y)[:, :g.OH, :g.OW].copy_(r)


def _w_from_dgrad_layout(g, wd):
    TA, TB = g.KH // g.S, g.KW // g.S
    t = wd.reshape(-1)[:g.S * g.S * g.Cin * TA * TB * g.Cout].view(g.S, g.S, g.Cin, TA, TB, g.Cout)  # py,px,c,a,b',n
    return t.permute(5, 2, 3, 0, 4, 1).reshape(g.Cout, g.Cin, g.KH, g.KW)                           # n,c,(a,py),(b',px)


def conv_dgrad(geom, dy, wd, dx, mask_src=None, slope=0.2, mask_bits=None, dbias_in=None, dbias_samples=0):
    g = geom
    w = _w_from_dgrad_layout(g, wd)
    d = _out_view(g, dy)[:, :g.OH, :g.OW].permute(0, 3, 1, 2)
    r = F.conv_transpose2d(d, w, stride=g.S)                                   # [B,Cin,(OH-1)S+KH,...]
    full = torch.zeros(g.B, g.Cin, g.H, g.W)
    full[:, :, :r.shape[2], :r.shape[3]] = r
    full = full.permute(0, 2, 3, 1)
    if mask_src is not None:
        full = full * _slope_mask(_in_view(g, mask_src)[:, :g.H, :g.W], slope)
    _in_view(g, dx)[:, :g.H, :g.W].copy_(full)
    if dbias_in is not None:
        nb = dbias_samples if dbias_samples > 0 else g.B
        dbias_in.view(-1)[:g.Cin].add_(full[:nb].sum((0, 1, 2)))


def conv_wgrad_splits(geom) -> int:
    return 2


def conv_wgrad(geom, dy, x, dw_partial, splits):
    g = geom
    xin = _in_view(g, x)[:, :g.H, :g.W].permute(0, 3, 1, 2)
    d = _out_view(g, dy)[:, :g.OH, :g.OW].permute(0, 3, 1, 2)
    with torch.enable_grad():
        w = torch.zeros(g.Cout, g.Cin, g.KH, g.KW, requires_grad=True)
        (F.conv2d(xin, w, None, stride=g.S)[:, :, :g.OH, :g.OW] * d).sum().backward()
    n = g.Cout * g.KH * g.KW * g.Cin
    p = dw_partial.view(-1)[:splits * n].view(splits, n)
    p.zero_()
    p[0] = w.grad.permute(0, 2, 3, 1).reshape(-1)


def linear_fwd(x, ldx, w, ldw, bias, y, ldy, M, N, K, epilogue, slope=0.2, splits=1):
    r = _v2(x, M, K, ldx) @ _v2(w, N, K, ldw).t()
    if epilogue in (EPI_BIAS_LRELU, EPI_BIAS):
        r = r + bias.view(-1)[:N]
    if epilogue == EPI_BIAS_LRELU:
        r = F.leaky_relu(r, slope)
    o = _v(y, (splits, M, N), (M * ldy, ldy, 1))
    o.zero_()
    o[0] = r


def linear_dgrad(dy, lddy, w, ldw, dx, lddx, M, N, K, mask_src=None, ldm=0, slope=0.2, mask_bits=None, colsum=None, colsum_mod=0,
                 colsum_rows=0):
    r = _v2(dy, M, K, lddy) @ _v2(w, K, N, ldw)
    if mask_src is not None:
        r = r * _slope_mask(_v2(mask_src, M, N, ldm), slope)
    _v2(dx, M, N, lddx).copy_(r)
    if colsum is not None:
        nr = colsum_rows if colsum_rows > 0 else M
        colsum.view(-1)[:colsum_mod].add_(r[:nr].reshape(nr, N // colsum_mod, colsum_mod).sum((0, 1)))


def linear_wgrad(dy, lddy, x, ldx, dw, lddw, M, N, K, splits=1):
    o = _v(dw, (splits, M, N), (M * lddw, lddw, 1))
    o.zero_()
    o[0] = _v2(dy, K, M, lddy).t() @ _v2(x, K, N, ldx)
