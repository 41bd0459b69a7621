"""Micro-benchmark: cost of the EPI_MASK epilogue (TMA-loaded LeakyReLU' source) vs plain epilogues, conv2/conv3 shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gail_carla_b200 import _abi as A, engine as E

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B = 4096
for layer in (2, 3):
    g = E.conv_geom(layer, B)
    cin, cout = E.CONV_CH[layer - 1], E.CONV_CH[layer]
    x = torch.randn(B * g.in_batch_stride, device="cuda"); y = torch.randn(B * g.out_batch_stride, device="cuda")
    dx = torch.zeros_like(x); w = torch.randn(cout * cin * 16, device="cuda") * 0.05; bias = torch.zeros(cout, device="cuda")
    fl = 2.0 * B * g.OH * g.OW * cout * 16 * cin
    for env in ({}, {"GC_NO_PATCH": "1"}):
        os.environ.pop("GC_NO_PATCH", None); os.environ.update(env)
        tag = "nopatch" if env else "patch"
        a = t(lambda: A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2))
        b = t(lambda: A.conv_fprop(g, x, w, None, y, A.EPI_MASK, 0.2, mask_src=y))
        c = t(lambda: A.conv_dgrad(g, y, w, dx, None, 0.2))
        d = t(lambda: A.conv_dgrad(g, y, w, dx, x, 0.2))
        xb = torch.randint(-2**31, 2**31 - 1, (x.numel() // 32,), dtype=torch.int32, device="cuda")
        e = t(lambda: A.conv_dgrad(g, y, w, dx, x, 0.2, mask_bits=xb))
        print(f"conv{layer} B={B} [{tag}] fprop bias+lrelu {a:.2f} ms ({fl/a/1e9:.0f} TF/s) | fprop masked {b:.2f} ms | dgrad plain {c:.2f} ms ({fl/c/1e9:.0f} TF/s) | dgrad masked {d:.2f} ms | dgrad bitmask {e:.2f} ms", flush=True)
    os.environ.pop("GC_NO_PATCH", None)
T, N = 4096, 18944
r = torch.rand(T, N, 1, device="cuda"); v = torch.randn(T + 1, N, 1, device="cuda"); m = torch.ones(T + 1, N, 1, device="cuda"); ret = torch.zeros_like(v)
fl_ = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(10):
    fl_.zero_(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); A.gae_returns(r, v, m, ret, 0.99, 0.95); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
tt = sorted(ts[3:])[len(ts[3:]) // 2] * 1e-3
print(f"GAE T={T} N={N}: {tt*1e6:.1f} us {16.0*T*N/tt/1e9:.0f} GB/s")
