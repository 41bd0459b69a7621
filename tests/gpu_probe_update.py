"""Diagnostic: run the full update iteration on the GPU for a golden case and print every deviation from the
reference's outputs (used to calibrate the TF32 tolerances asserted in tests/test_update_gpu.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import test_host_cpu as T

for name in sys.argv[1:]:
    t0 = time.time()
    z, pol, disc, ro, d_out, p_out, cl0, cl1 = T.run_update_case(name, "cuda")
    torch.cuda.synchronize()
    print(f"=== {name} ({time.time()-t0:.1f}s)")
    def rel(a, b):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        return float(np.nanmax(np.abs(a - b) / (1e-6 + np.abs(b))))
    print(" bootstrap value rel", rel(ro.value_preds[-1].cpu().numpy(), z["bootstrap_value"]))
    print(" compute_loss before", cl0, z["compute_loss_before"])
    print(" disc tuple", np.array(d_out), "\n        ref", z["disc_update"])
    print(" compute_loss after", cl1, z["compute_loss_after"])
    print(" gail_rewards rel", rel(ro.gail_rewards.cpu().numpy(), z["gail_rewards"]), " returns rel", rel(ro.returns.cpu().numpy()[:-1], z["returns"][:-1]))
    print(" ppo tuple", np.array([np.nan if x is None else x for x in p_out]), "\n       ref", z["ppo_update"])
    for prefix, sd in (("pol", pol.state_dict()), ("disc", disc.state_dict())):
        for k, v in sd.items():
            v = v.detach().float().reshape(-1).cpu()
            if v.numel() <= 4096: ref = z[f"{prefix}|{k}|full"]; got = v.numpy()
            else:
                stride = v.numel() // 2048; ref = z[f"{prefix}|{k}|sample"]; got = v[::stride][:2048].numpy()
            d = np.abs(got - ref)
            print(f"  {prefix} {k:52s} max {d.max():.2e} mean {d.mean():.2e} frac>2e-5 {(d > 2e-5).mean():.3f}")
