"""Micro-driver for ncu: a few launches of the main conv contractions at B=2048 (conv1..3 fprop, dgrad, wgrad)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gail_carla_b200 import _abi as A, engine as E
B = int(os.environ.get("B", 2048))
reps = int(os.environ.get("REPS", 2))
def t(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for layer in (1, 2, 3):
    g = E.conv_geom(layer, B)
    cin, cout = E.CONV_CH[layer - 1], E.CONV_CH[layer]
    nw = 2048 if layer == 1 else cout * cin * 16
    x = torch.randn(B * g.in_batch_stride, device="cuda"); y = torch.randn(B * g.out_batch_stride, device="cuda")
    dx = torch.zeros_like(x); w = torch.randn(nw, device="cuda") * 0.05; bias = torch.zeros(cout, device="cuda")
    bits = torch.zeros(y.numel() // 32, dtype=torch.int32, device="cuda")
    xbits = torch.zeros(max(1, x.numel() // 32), dtype=torch.int32, device="cuda")
    fl = 2.0 * B * g.OH * g.OW * cout * g.KH * g.KW * g.Cin
    a = t(lambda: A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2))
    b = t(lambda: A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2, mask_bits=bits))
    c = t(lambda: A.conv_dgrad(g, y, w, dx, None, 0.2))
    d = t(lambda: A.conv_dgrad(g, y, w, dx, x, 0.2)) if layer > 1 else float("nan")
    e = t(lambda: A.conv_dgrad(g, y, w, dx, x, 0.2, mask_bits=xbits)) if layer > 1 else float("nan")
    z = A.conv_wgrad_splits(g); part = torch.zeros(z * nw, device="cuda")
    f = t(lambda: A.conv_wgrad(g, y, x, part, z))
    print(f"conv{layer} B={B}: fprop {a:.2f} ms ({fl/a/1e9:.0f} TF/s) | fprop+bits {b:.2f} | dgrad {c:.2f} ({fl/c/1e9:.0f}) | dgrad tma-mask {d:.2f} | dgrad bitmask {e:.2f} | wgrad {f:.2f} ({fl/f/1e9:.0f}) splits={z}", flush=True)
