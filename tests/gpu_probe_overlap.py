"""Diagnostic: does a pinned H2D copy on a side stream overlap with our persistent GEMM kernels on the main stream?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gail_carla_b200 import _abi as A, engine as E
B = 4096
g = E.conv_geom(3, B)
x = torch.randn(B * g.in_batch_stride, device="cuda"); y = torch.zeros(B * g.out_batch_stride, device="cuda")
w = torch.randn(128 * 64 * 16, device="cuda") * 0.05; bias = torch.zeros(128, device="cuda")
host = torch.empty(B * 3 * 192 * 192, pin_memory=True); devb = torch.empty_like(host, device="cuda")
side = torch.cuda.Stream()
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")

def compute_ours(n=40):
    for _ in range(n): A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2)
def compute_torch(n=12):
    for _ in range(n): torch.matmul(a, b)
def copy():
    with torch.cuda.stream(side): devb.copy_(host, non_blocking=True)
def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for name, comp in (("ours", compute_ours), ("torch matmul", compute_torch)):
    comp(2); torch.cuda.synchronize()
    tc = wall(comp); tp = wall(copy)
    tb = wall(lambda: (copy(), comp()))
    tb2 = wall(lambda: (comp(), copy()))
    print(f"[{name}] compute {tc:.1f} ms | copy {tp:.1f} ms | copy-then-compute enqueued {tb:.1f} ms | compute-then-copy enqueued {tb2:.1f} ms", flush=True)
