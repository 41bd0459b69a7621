"""Diagnostic: timeline of the H2D expert copies against the real Discriminator.update batches."""
import os, sys, time, torch
from types import SimpleNamespace as NS
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gail_carla_b200 as G
from gail_carla_b200 import synthetic
dev = torch.device("cuda", 0)
T, N, B = 256, 64, 4096
torch.manual_seed(1)
disc = G.Discriminator(synthetic.OBS_SHAPE, NS(shape=(4,)), NS(shape=(2,)), 100, dev, 2.5e-4, 1e-8, (0.9, 0.99), 0.5).to(dev)
ro = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device=dev)
synthetic.fill_rollout(ro, seed=11, chunk=64)
loader = synthetic.SyntheticExpertLoader(T * N // B, B, seed=21, pin=True)
disc.update(loader, ro); torch.cuda.synchronize()
orig = disc.engine.update_step
marks = []
def traced(*a, **k):
    s = torch.cuda.Event(enable_timing=True); s.record()
    r = orig(*a, **k)
    e = torch.cuda.Event(enable_timing=True); e.record()
    marks.append((s, e, time.perf_counter()))
    return r
disc.engine.update_step = traced
for rep in range(2):
    marks.clear(); disc._trace = []
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record(); h0 = time.perf_counter()
    disc.update(loader, ro)
    torch.cuda.synchronize()
    print(f"rep {rep}: total {(time.perf_counter()-h0)*1e3:.0f} ms")
    for i, (tag, c0, c1) in enumerate(disc._trace): print(f"   copy {i}: {t0.elapsed_time(c0):.0f} -> {t0.elapsed_time(c1):.0f} ms")
    for i, (s, e, h) in enumerate(marks): print(f"   update_step {i}: {t0.elapsed_time(s):.0f} -> {t0.elapsed_time(e):.0f} ms (host enqueued at {(h-h0)*1e3:.0f})")
