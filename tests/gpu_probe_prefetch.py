"""Diagnostic: timeline of Discriminator._prefetched (copy stream) against a fake 100 ms consumer on the main stream."""
import os, sys, time, torch
from types import SimpleNamespace as NS
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gail_carla_b200 as G
from gail_carla_b200 import synthetic
dev = torch.device("cuda", 0)
disc = G.Discriminator(synthetic.OBS_SHAPE, NS(shape=(4,)), NS(shape=(2,)), 100, dev, 2.5e-4, 1e-8, (0.9, 0.99), 0.5).to(dev)
disc.engine.sync_params()
B, nb = 4096, 5
loader = synthetic.SyntheticExpertLoader(nb, B, seed=21, pin=True)
print("pinned:", all(t.is_pinned() for b in loader for t in b))
cycles = int(0.1 * 1.9e9)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    marks = []
    h0 = time.perf_counter()
    for i, (ts, idx, release) in enumerate(disc._prefetched(zip(loader, range(nb)))):
        s = torch.cuda.Event(enable_timing=True); s.record()
        x = ts[0].sum()          # consume
        release()
        torch.cuda._sleep(cycles)
        e = torch.cuda.Event(enable_timing=True); e.record()
        marks.append((s, e, (time.perf_counter() - h0) * 1e3))
    torch.cuda.synchronize()
    print(f"rep {rep}: " + " | ".join(f"b{i}: start {t0.elapsed_time(s):.0f} end {t0.elapsed_time(e):.0f} (host enq {h:.0f})" for i, (s, e, h) in enumerate(marks)))
