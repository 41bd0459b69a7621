"""GPU bring-up probe for the tcgen05 GEMM engine: every case compares a C-ABI contraction with its CPU statement in
oracle/abi_emu.py on small-integer inputs (exact in TF32, so any mismatch is a layout / descriptor bug, not rounding).
Each case runs in its own subprocess so a trapped kernel cannot poison the others.

    python tests/gpu_probe_umma.py            # all cases -> gpurun_out/probe_umma.log
    python tests/gpu_probe_umma.py --case NAME
"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ints(shape, lo=-2, hi=3, seed=0):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi, shape, generator=g).float()


def report(name, got, ref):
    import torch
    got = got.float().cpu(); ref = ref.float().cpu()
    diff = (got - ref).abs()
    bad = diff > 1e-3 * (1 + ref.abs())
    nbad = int(bad.sum())
    rel = float((got - ref).norm() / (ref.norm() + 1e-20))
    print(f"[{name}] shape={tuple(ref.shape)} mismatches={nbad}/{ref.numel()} max_abs={float(diff.max()):.4g} rel_fro={rel:.3g} "
          f"ref_norm={float(ref.norm()):.4g} got_norm={float(got.norm()):.4g}", flush=True)
    if nbad:
        idx = bad.nonzero()[:12]
        for i in idx:
            t = tuple(int(v) for v in i)
            print(f"    at {t}: got {float(got[t]):.4g} ref {float(ref[t]):.4g}")
        # which rows / cols are affected
        if ref.dim() == 2:
            rows = bad.any(1).nonzero().view(-1)[:40].tolist(); cols = bad.any(0).nonzero().view(-1)[:40].tolist()
            print(f"    bad rows (first 40): {rows}\n    bad cols (first 40): {cols}")
    return nbad == 0


def geom(layer, B):
    from gail_carla_b200._abi import ConvGeom, LDF
    if layer == 1:   # conv1 in space-to-depth form: 2x2 taps, stride 1, 16 -> 32, 96x96 -> 95x95 (pitch 96)
        return ConvGeom(B, 96, 96, 96, 96, 16, 2, 2, 1, 95, 95, 96, 96, 32, 96 * 96 * 16, 96 * 96 * 32)
    if layer == 2:
        return ConvGeom(B, 95, 95, 96, 96, 32, 4, 4, 2, 46, 46, 46, 46, 64, 96 * 96 * 32, 46 * 46 * 64)
    if layer == 3:
        return ConvGeom(B, 46, 46, 46, 46, 64, 4, 4, 2, 22, 22, 22, 22, 128, 46 * 46 * 64, 22 * 22 * 128)
    if layer == 4:   # output lands in the feature matrix (row pitch LDF)
        return ConvGeom(B, 22, 22, 22, 22, 128, 4, 4, 2, 10, 10, 10, 10, 256, 22 * 22 * 128, LDF)
    raise ValueError(layer)


def run_case(name):
    import torch
    from gail_carla_b200 import _abi as A
    from oracle import abi_emu as E
    dev = "cuda"
    ok = True
    t0 = time.time()
    kind, *rest = name.split(":")
    if kind == "lin":
        M, N, K, epi, splits = (int(v) for v in rest)
        ldx, ldw, ldy = K + (4 - K % 4) % 4, K + (4 - K % 4) % 4, N + (4 - N % 4) % 4
        x = torch.zeros(M, ldx); x[:, :K] = ints((M, K), seed=1)
        w = torch.zeros(N, ldw); w[:, :K] = ints((N, K), seed=2)
        b = ints((N,), seed=3)
        y_ref = torch.zeros(splits, M, ldy); y = torch.full((splits, M, ldy), 7.0, device=dev)
        E.linear_fwd(x, ldx, w, ldw, b, y_ref, ldy, M, N, K, epi, 0.5, splits)
        A.linear_fwd(x.to(dev), ldx, w.to(dev), ldw, b.to(dev), y, ldy, M, N, K, epi, 0.5, splits)
        torch.cuda.synchronize()
        ok = report(name, y.sum(0)[:, :N], y_ref.sum(0)[:, :N])
    elif kind == "lind":   # dgrad
        M, N, K, masked = (int(v) for v in rest)
        dy = ints((M, K), seed=1); w = ints((K, N), seed=2); ms = ints((M, N), seed=3)
        dx_ref = torch.zeros(M, N); dx = torch.full((M, N), 7.0, device=dev)
        E.linear_dgrad(dy, K, w, N, dx_ref, N, M, N, K, ms if masked else None, N, 0.5)
        A.linear_dgrad(dy.to(dev), K, w.to(dev), N, dx, N, M, N, K, ms.to(dev) if masked else None, N, 0.5)
        torch.cuda.synchronize()
        ok = report(name, dx, dx_ref)
    elif kind == "linw":   # wgrad
        M, N, K, splits = (int(v) for v in rest)
        dy = ints((K, M), seed=1); x = ints((K, N), seed=2)
        dw_ref = torch.zeros(splits, M, N); dw = torch.full((splits, M, N), 7.0, device=dev)
        E.linear_wgrad(dy, M, x, N, dw_ref, N, M, N, K, splits)
        A.linear_wgrad(dy.to(dev), M, x.to(dev), N, dw, N, M, N, K, splits)
        torch.cuda.synchronize()
        ok = report(name, dw.sum(0), dw_ref.sum(0))
    elif kind in ("cf", "cd", "cw"):
        layer, B, flag = (int(v) for v in rest)
        g = geom(layer, B)
        nin = B * g.in_batch_stride; nout = B * g.out_batch_stride
        Kw = g.KH * g.KW * g.Cin
        x = ints((nin,), seed=1); dy = ints((nout,), seed=2)
        bias = ints((g.Cout,), seed=4)
        if kind == "cf":
            w = ints((g.Cout * Kw,), -1, 2, seed=3)
            ms = ints((nout,), seed=5)
            y_ref = torch.zeros(nout); y = torch.zeros(nout, device=dev)
            epi = flag
            E.conv_fprop(g, x, w, bias, y_ref, epi, 0.5, ms if epi == 3 else None)
            A.conv_fprop(g, x.to(dev), w.to(dev), bias.to(dev), y, epi, 0.5, ms.to(dev) if epi == 3 else None)
            torch.cuda.synchronize()
            ok = report(name, E._out_view(g, y.cpu())[:, :g.OH, :g.OW].reshape(-1, g.Cout),
                        E._out_view(g, y_ref)[:, :g.OH, :g.OW].reshape(-1, g.Cout))
        elif kind == "cd":
            wd = ints((g.Cout * Kw,), -1, 2, seed=3)
            ms = ints((nin,), seed=5)
            dx_ref = torch.zeros(nin); dx = torch.zeros(nin, device=dev)
            E.conv_dgrad(g, dy, wd, dx_ref, ms if flag else None, 0.5)
            A.conv_dgrad(g, dy.to(dev), wd.to(dev), dx, ms.to(dev) if flag else None, 0.5)
            torch.cuda.synchronize()
            ok = report(name, E._in_view(g, dx.cpu())[:, :g.H, :g.W].reshape(-1, g.Cin),
                        E._in_view(g, dx_ref)[:, :g.H, :g.W].reshape(-1, g.Cin))
        else:
            splits = A.conv_wgrad_splits(g) if flag == 0 else flag
            print(f"    wgrad splits={splits}")
            # zero the padding of dy so both sides agree on what padding pixels hold
            dyv = E._out_view(g, dy); keep = dyv[:, :g.OH, :g.OW].clone(); dy.zero_(); dyv[:, :g.OH, :g.OW] = keep
            p_ref = torch.zeros(splits, g.Cout * Kw); p = torch.full((splits, g.Cout * Kw), 7.0, device=dev)
            E.conv_wgrad(g, dy, x, p_ref, splits)
            A.conv_wgrad(g, dy.to(dev), x.to(dev), p, splits)
            torch.cuda.synchronize()
            ok = report(name, p.sum(0).view(g.Cout, Kw), p_ref.sum(0).view(g.Cout, Kw))
    else:
        raise ValueError(name)
    print(f"    {name}: {'OK' if ok else 'FAIL'} in {time.time() - t0:.1f}s", flush=True)
    return ok


CASES = [
    "lin:128:128:32:0:1", "lin:128:128:128:0:1", "lin:256:256:512:1:1", "lin:200:100:1000:2:1", "lin:300:512:2048:0:4",
    "lin:64:16:96:0:1", "lin:512:100:25632:0:8",
    "lind:128:256:64:0", "lind:300:512:100:1", "lind:256:25600:512:0",
    "linw:128:256:128:1", "linw:100:25632:384:2", "linw:512:512:200:1",
    "cf:2:2:1", "cf:3:3:1", "cf:4:7:1", "cf:1:2:1", "cf:2:2:3", "cf:1:2:3",
    "cd:2:2:0", "cd:3:2:1", "cd:4:7:1", "cd:1:2:0",
    "cw:2:2:0", "cw:3:3:0", "cw:4:8:0", "cw:1:2:0", "cw:4:7:3",
]


def main():
    if "--case" in sys.argv:
        ok = run_case(sys.argv[sys.argv.index("--case") + 1])
        sys.exit(0 if ok else 1)
    cases = CASES
    if "--only" in sys.argv:
        pref = sys.argv[sys.argv.index("--only") + 1].split(",")
        cases = [c for c in CASES if any(c.startswith(p) for p in pref)]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "probe_umma.log"), "w")
    n_ok = 0
    for c in cases:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", c], capture_output=True, text=True,
                               timeout=180)
            out = r.stdout + ("\n" + r.stderr[-1500:] if r.returncode not in (0, 1) or "Error" in r.stderr else "")
            status = r.returncode
        except subprocess.TimeoutExpired as e:
            out, status = f"[{c}] TIMEOUT\n{(e.stdout or b'').decode()[-500:]}", -9
        n_ok += status == 0
        msg = f"=== {c} -> exit {status}\n{out}\n"
        log.write(msg); log.flush()
        print(msg, flush=True)
    summary = f"SUMMARY: {n_ok}/{len(cases)} cases OK"
    log.write(summary + "\n"); print(summary)


if __name__ == "__main__":
    main()
