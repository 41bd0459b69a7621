"""Calibration probe (not a test): prints the observed deviations of the CUDA path from the reference goldens and from
the oracle's autograd gradients, so the tolerances in tests/test_update_gpu.py / test_grads_gpu.py can be set from
measurements.  Run on the GPU box: python tests/gpu_probe_parity.py"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import test_host_cpu as H
import grad_cases as GC


def report_case(name, **kw):
    t0 = time.time()
    z, pol, disc, ro, d_out, p_out, cl0, cl1 = H.run_update_case(name, "cuda", **kw)
    def rel(a, b):
        a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
        return float(np.nanmax(np.abs(a - b) / (1e-3 + np.abs(b))))
    out = dict(bootstrap=rel(ro.value_preds[-1].cpu().numpy(), z["bootstrap_value"]), cl0=rel(cl0, z["compute_loss_before"]),
               disc=rel(d_out, z["disc_update"]), cl1=rel(cl1, z["compute_loss_after"]),
               rewards=rel(ro.gail_rewards.cpu().numpy(), z["gail_rewards"]), returns=rel(ro.returns.cpu().numpy(), z["returns"]),
               ppo=rel([np.nan if x is None else x for x in p_out], z["ppo_update"]))
    worst = {}
    for prefix, mod, lr in (("disc", disc, H.HP["gail_lr"]), ("pol", pol, H.HP["lr"])):
        mx = mn = 0.0
        for k, v in mod.state_dict().items():
            v = v.detach().float().reshape(-1).cpu()
            if v.numel() <= 4096:
                ref = z[f"{prefix}|{k}|full"]; got = v.numpy()
            else:
                stride = v.numel() // 2048
                ref = z[f"{prefix}|{k}|sample"]; got = v[::stride][:2048].numpy()
            d = np.abs(got - ref)
            mx = max(mx, float(d.max()) / lr); mn = max(mn, float(d.mean()) / lr)
        worst[prefix] = (round(mx, 3), round(mn, 4))
    print(f"{name} {kw}: scaled errs {{{', '.join(f'{k}: {v:.2e}' for k, v in out.items())}}} params max|d|/lr, mean|d|/lr: {worst} ({time.time()-t0:.1f}s)", flush=True)


def report_grads():
    def worst(got, ref):
        wc, wr, wk = 1.0, 0.0, None
        per = {}
        for k, r in ref.items():
            g = got[k].double().reshape(-1); r = r.double().reshape(-1)
            if r.norm() == 0:
                continue
            cos = float(torch.dot(g, r) / (r.norm() * g.norm())); rel = float((g - r).norm() / r.norm())
            per[k] = rel
            if rel > wr:
                wc, wr, wk = cos, rel, k
        return wc, wr, wk, per
    for B, Be in ((64, 0), (48, 16), (200, 0)):
        got, ref = GC.policy_grads("cuda", B, Be)
        wc, wr, wk, per = worst(got, ref)
        print(f"policy grads B={B} Be={Be}: worst rel-Fro {wr:.2e} (cos {wc:.6f}) at {wk}; conv rel: " +
              ", ".join(f"{k.split('main.')[1]}={v:.1e}" for k, v in per.items() if "main" in k), flush=True)
    for B in (32, 100):
        got, ref, gs, rs = GC.critic_grads("cuda", B)
        wc, wr, wk, per = worst(got, ref)
        print(f"critic grads B={B}: worst rel-Fro {wr:.2e} (cos {wc:.6f}) at {wk}; wd {gs['wd']:.6f}/{rs['wd']:.6f} gp {gs['gp']:.6f}/{rs['gp']:.6f}; conv rel: " +
              ", ".join(f"{k.split('main.')[1]}={v:.1e}" for k, v in per.items() if "main" in k), flush=True)


def report_extra():
    for B in (64, 200):
        tf32, fp32 = GC.stock_tf32_policy_grads(B)
        e = GC.rel_errors(tf32, fp32)
        k = max(e, key=lambda n: e[n][1])
        print(f"stock PyTorch TF32 (cuDNN/cuBLAS) vs CPU fp32, policy grads B={B}: worst rel-Fro {e[k][1]:.2e} (cos {e[k][0]:.6f}) at {k}; conv rel: " +
              ", ".join(f"{n.split('main.')[1]}={v[1]:.1e}" for n, v in e.items() if "main" in n), flush=True)
        got, ref = GC.policy_grads("cuda", B, 0)
        m = GC.rel_errors(got, ref)
        print(f"   this repo vs CPU fp32, same minibatch: " + ", ".join(f"{n.split('main.')[1]}={v[1]:.1e}" for n, v in m.items() if "main" in n)
              + " | fc: " + ", ".join(f"{v[1]:.1e}" for n, v in m.items() if "main" not in n), flush=True)
        print(f"   ratio mine/stock per tensor: " + ", ".join(f"{m[n][1] / max(e[n][1], 1e-12):.2f}" for n in m), flush=True)
    with GC.linearised():
        for B, Be in ((64, 0), (200, 0)):
            got, ref = GC.policy_grads("cuda", B, Be)
            e = GC.rel_errors(got, ref)
            k = max(e, key=lambda n: e[n][1])
            print(f"linearised (slope 1) policy grads B={B}: worst rel-Fro {e[k][1]:.2e} (cos {e[k][0]:.7f}) at {k}", flush=True)
        for B in (32, 100):
            got, ref, gs, rs = GC.critic_grads("cuda", B)
            e = GC.rel_errors(got, ref)
            k = max(e, key=lambda n: e[n][1])
            print(f"linearised (slope 1) critic grads B={B}: worst rel-Fro {e[k][1]:.2e} (cos {e[k][0]:.7f}) at {k}; gp {gs['gp']:.6f}/{rs['gp']:.6f}", flush=True)


if __name__ == "__main__":
    if "--extra" in sys.argv:
        report_extra()
        sys.exit(0)
    report_grads()
    for name in ("update_tiny", "update_tiny2", "update_c1", "update_unclipped", "update_mid"):
        report_case(name)
    report_case("update_tiny", obs_dtype=torch.uint8, expert_u8=True)
    report_case("update_mid", obs_dtype=torch.uint8, expert_u8=True)
