"""Micro-driver for ncu: conv1 fprop (space-to-depth 16->32) with the LeakyReLU' bit mask at B=4096, three launches."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gail_carla_b200 import _abi as A, engine as E
B = int(os.environ.get("B", 4096))
g = E.conv_geom(1, B)
x = torch.randn(B * g.in_batch_stride, device="cuda"); y = torch.empty(B * g.out_batch_stride, device="cuda")
w = torch.randn(2048, device="cuda") * 0.05; bias = torch.zeros(32, device="cuda")
bits = torch.zeros(y.numel() // 32, dtype=torch.int32, device="cuda")
for _ in range(3):
    A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2, mask_bits=bits)
torch.cuda.synchronize()
print("ok")
