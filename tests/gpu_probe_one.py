"""Micro-driver for ncu: OP=fprop|dgrad|wgrad LAYER=1..4 B=... launches that one contraction REPS times."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gail_carla_b200 import _abi as A, engine as E
B = int(os.environ.get("B", 2048)); layer = int(os.environ.get("LAYER", 1)); op = os.environ.get("OP", "fprop")
reps = int(os.environ.get("REPS", 3))
g = E.conv_geom(layer, B)
cin, cout = E.CONV_CH[layer - 1], E.CONV_CH[layer]
nw = 2048 if layer == 1 else cout * cin * 16
x = torch.randn(B * g.in_batch_stride, device="cuda"); y = torch.randn(B * g.out_batch_stride, device="cuda")
dx = torch.zeros_like(x); w = torch.randn(nw, device="cuda") * 0.05; bias = torch.zeros(cout, device="cuda")
z = A.conv_wgrad_splits(g); part = torch.zeros(z * nw, device="cuda")
bits = torch.zeros(y.numel() // 32, dtype=torch.int32, device="cuda") if os.environ.get("BITS") else None
fn = {"fprop": lambda: A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2, mask_bits=bits),
      "dgrad": lambda: A.conv_dgrad(g, y, w, dx, None, 0.2),
      "wgrad": lambda: A.conv_wgrad(g, y, x, part, z)}[op]
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("ok", op, layer, B)
