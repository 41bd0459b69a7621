"""Diagnostic: phase breakdown of one update iteration (c4-like, fewer steps) with host-pinned vs device-resident expert data."""
import os, sys, time, torch
from types import SimpleNamespace as NS
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gail_carla_b200 as G
from gail_carla_b200 import synthetic
sys.path.insert(0, ROOT)
import bench as Bn
HP = Bn.HP
dev = torch.device("cuda", 0)
T, N, B = int(os.environ.get("T", 256)), 64, 4096
sp, asp = NS(shape=(4,)), NS(shape=(2,))
torch.manual_seed(1)
pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False).to(dev)
agent = G.PPO(pol, 0.1, 1, B, 0.5, dev, lr=1e-4, eps=1e-8, betas=(0.9, 0.99), max_grad_norm=0.5)
disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, dev, 2.5e-4, 1e-8, (0.9, 0.99), 0.5).to(dev)
ro = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device=dev)
synthetic.fill_rollout(ro, seed=11, chunk=64)
nb = T * N // B
host_loader = synthetic.SyntheticExpertLoader(nb, B, seed=21, pin=True)
class DevLoader:
    def __init__(self, l): self.batch_size = l.batch_size; self.b = [tuple(t.to(dev) for t in x) for x in l]
    def __len__(self): return len(self.b)
    def __iter__(self): return iter(self.b)
dev_loader = DevLoader(host_loader)

def phase(fn):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), (t1 - t0) * 1e3

for name, loader in (("host-pinned expert", host_loader), ("device expert", dev_loader), ("host-pinned expert", host_loader)):
    for rep in range(2):
        r = {}
        r["get_value"] = phase(lambda: ro.value_preds[-1].copy_(pol.get_value(ro.obs[-1], ro.metrics[-1])))
        r["disc.update"] = phase(lambda: disc.update(loader, ro))
        r["predict_rewards"] = phase(lambda: disc.predict_rewards_rollout(ro))
        r["compute_returns"] = phase(lambda: ro.compute_returns(0.99, 0.95))
        r["ppo.update"] = phase(lambda: agent.update(ro))
        r["after_update"] = phase(lambda: ro.after_update())
    tot = sum(v[0] for v in r.values())
    print(f"[{name}] T={T}: total {tot:.1f} ms | " + " | ".join(f"{k} gpu {v[0]:.1f} ms (host {v[1]:.1f})" for k, v in r.items()), flush=True)
