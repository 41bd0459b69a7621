#!/usr/bin/env python
"""Benchmark of the gail-carla learning hot path on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c4|c2|c5|tiny]

A *step* is one full PPO+WDGAIL update on a collected rollout (tools/learn.py:137-223,269 minus diagnostics):
bootstrap value -> Discriminator.update x gail_epoch -> predict_reward over all T*N -> compute_returns (GAE) ->
PPO.update -> after_update.  Metric: env-steps/s = T*N / step time, whole job over all ranks (strong scaling: the
rollout and the minibatch are split across ranks by environment, gradients are all-reduced with NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace as NS

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HP = dict(lr=1e-4, eps=1e-8, betas=(0.9, 0.99), clip_param=0.1, value_loss_coef=0.5, max_grad_norm=0.5,
          gail_lr=2.5e-4, gail_eps=1e-8, gail_betas=(0.9, 0.99), gail_max_grad_norm=0.5, gamma=0.99, gae_lambda=0.95,
          logstd=[-1.4, -3.2])   # params_variable.json:29-55

CONFIGS = {
    # BASELINE.json configs[3]: the configuration the env-steps/s metric is quoted on
    "c4": dict(T=1024, N=64, B_ppo=4096, B_gail=4096, ppo_epoch=1, gail_epoch=1,
               workload="configs[3]: full PPO+WDGAIL update epoch, 64 envs x 1024 steps, B=4096"),
    "c5": dict(T=512, N=256, B_ppo=4096, B_gail=4096, ppo_epoch=10, gail_epoch=1,
               workload="configs[4]: 256 envs x 512 steps, 10 PPO epochs x 32 minibatches"),
    # same batch shapes as c4 with 1/8 of the rollout: used for the ncu launch list (profiles/)
    "c4s": dict(T=128, N=64, B_ppo=4096, B_gail=4096, ppo_epoch=1, gail_epoch=1,
                workload="c4 batch shapes on a 64 envs x 128 steps rollout (profiling)"),
    # the per-rank share of c4 at 8 GPUs (8 envs, 512-row minibatches) on one GPU: isolates launch / host overheads
    "c4r8": dict(T=1024, N=8, B_ppo=512, B_gail=512, ppo_epoch=1, gail_epoch=1,
                 workload="per-rank share of configs[3] at 8 GPUs: 8 envs x 1024 steps, B=512"),
    "c4r4": dict(T=1024, N=16, B_ppo=1024, B_gail=1024, ppo_epoch=1, gail_epoch=1,
                 workload="per-rank share of configs[3] at 4 GPUs: 16 envs x 1024 steps, B=1024"),
    # stock-PyTorch-on-CUDA yardstick (`gpu_baseline`): a shorter rollout with the batch sizes autograd's graphs allow
    "g2048": dict(T=128, N=64, B_ppo=2048, B_gail=2048, ppo_epoch=1, gail_epoch=1, workload="64 envs x 128 steps, B=2048"),
    "g1024": dict(T=128, N=64, B_ppo=1024, B_gail=1024, ppo_epoch=1, gail_epoch=1, workload="64 envs x 128 steps, B=1024"),
    "g512": dict(T=128, N=64, B_ppo=512, B_gail=512, ppo_epoch=1, gail_epoch=1, workload="64 envs x 128 steps, B=512"),
    "tiny": dict(T=64, N=8, B_ppo=128, B_gail=128, ppo_epoch=1, gail_epoch=1, workload="tiny: 8 envs x 64 steps, B=128"),
    # the bounded sample the CPU arm runs: configs[0] verbatim
    "c1": dict(T=128, N=1, B_ppo=128, B_gail=128, ppo_epoch=1, gail_epoch=1,
               workload="configs[0]: 1 env x 128 steps, B=128"),
    "c1b512": dict(T=512, N=1, B_ppo=512, B_gail=512, ppo_epoch=1, gail_epoch=1,
                   workload="1 env x 512 steps, B=512 (batch-size bracket of the CPU sample)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops_sustained"], bf16_burst=d["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ reference arm
class _DeviceLoader:
    """Expert loader whose batches already live on `device` (yardstick runs only)."""
    def __init__(self, inner, device):
        self.batch_size = inner.batch_size
        self._b = [tuple(t.to(device) for t in b) for b in inner]
    def __len__(self):
        return len(self._b)
    def __iter__(self):
        return iter(self._b)


def oracle_update_seconds(c, device="cpu", repeats=1):
    """One full update (tools/learn.py:137-223,269 minus diagnostics) of config `c` through the oracle restatement -
    the reference's own stock-PyTorch code path, function for function - on `device`; returns seconds per update.
    device="cpu": torch CPU fp32 on all host threads (the reference arm / cpu_baseline).  device="cuda": the same
    PyTorch code on this GPU with cuDNN / cuBLAS free to use TF32 (`gpu_baseline`: the like-for-like competitor of the
    hand-written kernels; data resident on the device, nothing of this repo's package on that path)."""
    from gail_carla_b200 import synthetic
    from oracle import ref_path as O
    T, N = c["T"], c["N"]
    dev = torch.device(device)
    torch.manual_seed(1)
    pol, disc = O.init_policy_params(), O.init_disc_params()
    pol = {k: v.to(dev) for k, v in pol.items()}; disc = {k: v.to(dev) for k, v in disc.items()}
    padam = O.AdamState(pol, HP["lr"], HP["eps"], HP["betas"]); dadam = O.AdamState(disc, HP["gail_lr"], HP["gail_eps"], HP["gail_betas"])
    ro = NS(obs=torch.zeros(T + 1, N, 3, 192, 192, device=dev), metrics=torch.zeros(T + 1, N, 4, device=dev),
            actions=torch.zeros(T, N, 2, device=dev), action_log_probs=torch.zeros(T, N, 1, device=dev),
            value_preds=torch.zeros(T + 1, N, 1, device=dev), returns=torch.zeros(T + 1, N, 1, device=dev),
            masks=torch.ones(T + 1, N, 1, device=dev), gail_rewards=torch.zeros(T, N, 1, device=dev),
            rewards=torch.zeros(T, N, 1, device=dev), num_steps=T, num_processes=N)
    synthetic.fill_rollout(ro, seed=11, chunk=max(1, 2048 // N))
    loader = synthetic.SyntheticExpertLoader(T * N // c["B_gail"], c["B_gail"], seed=21)
    if dev.type == "cuda":
        loader = _DeviceLoader(loader, dev)
    d = {k: getattr(ro, k) for k in ("obs", "metrics", "actions", "action_log_probs", "value_preds", "returns", "masks",
                                    "gail_rewards", "rewards")}

    def one_update():
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            d["value_preds"][-1] = O.policy_base(pol, d["obs"][-1], d["metrics"][-1], True, HP["logstd"])[0]
        for _ in range(c["gail_epoch"]):
            O.disc_update(disc, dadam, loader, d, HP["gail_max_grad_norm"])
        for step in range(T):
            d["gail_rewards"][step] = O.predict_reward(disc, d["obs"][step], d["metrics"][step], d["actions"][step])
        d["returns"] = O.gae_returns(d["gail_rewards"], d["value_preds"], d["masks"], HP["gamma"], HP["gae_lambda"])
        O.ppo_update(pol, padam, d, clip_param=HP["clip_param"], ppo_epoch=c["ppo_epoch"], mini_batch_size=c["B_ppo"],
                     value_loss_coef=HP["value_loss_coef"], max_grad_norm=HP["max_grad_norm"], logstd=HP["logstd"])
        if dev.type == "cuda":
            torch.cuda.synchronize()
        return time.perf_counter() - t0

    if dev.type == "cuda":
        one_update()                      # warm-up: cuDNN algorithm selection, allocator growth
    return sum(one_update() for _ in range(repeats)) / repeats


def cpu_update_seconds(c):
    return oracle_update_seconds(c, "cpu")


def gpu_baseline_sample(dev):
    """Stock PyTorch on the same B200 (BASELINE.md section 4 "also reported where cheap"): the oracle restatement of the
    reference's update on CUDA tensors, cuDNN autotuned, TF32 allowed for convolutions and matmuls (the reference's own
    GPU path runs that way under its torch version).  Largest batch whose double-backward graph fits next to the
    resident rollout is tried first."""
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    out = None
    try:
        for name in ("g2048", "g1024", "g512"):
            c = CONFIGS[name]
            try:
                torch.cuda.empty_cache()
                t = oracle_update_seconds(c, dev, repeats=2)
                out = {"value": c["T"] * c["N"] / t, "unit": "env-steps/s", "seconds_per_update": t, "kind": "port",
                       "what": "oracle/ref_path.py (the reference's stock PyTorch update, function for function) on cuda: cuDNN "
                               "autotuned, TF32 allowed, rollout + expert batches resident on the device",
                       "sample": c["workload"] + "; env-steps/s is per-sample work, so the bounded sample stands for the full workload"}
                break
            except torch.cuda.OutOfMemoryError:
                out = {"unavailable": f"{name}: out of memory"}
                continue
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()
    return out


def run_reference(args):
    """The reference's CPU torch path for the same update, through the oracle restatement (oracle/ref_path.py, pinned
    against the unmodified reference - the reference tree itself does not exist on the GPU box), all host threads.
    A step is a BOUNDED SAMPLE of the workload: one full update of configs[0] (1 env x 128 steps, B=128) - the line's
    `config` names what actually ran (`config.workload`) and the workload it stands for (`config.sample_of`).  The
    per-sample cost of the CPU path falls with the batch size, so one extra update at B=512 (`cpu_baseline_b512`) brackets
    the extrapolation from B=128 to the GPU arm's B=4096."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    c = CONFIGS["c1"]
    T, N = c["T"], c["N"]
    for _ in range(args.warmup):
        cpu_update_seconds(c)
    times = [cpu_update_seconds(c) for _ in range(args.steps)]
    dt = sum(times) / len(times)
    v = T * N / dt
    cfg = CONFIGS[args.config]
    b512 = None
    if not args.no_b512:
        c512 = CONFIGS["c1b512"]
        t512 = cpu_update_seconds(c512)
        b512 = {"value": c512["T"] * c512["N"] / t512, "unit": "env-steps/s", "seconds": t512, "cores": torch.get_num_threads(),
                "kind": "port", "sample": "one full update on " + c512["workload"] + ", run once after the timed steps"}
    line = {"impl": "reference", "metric": "ppo_wdgail_update_env_steps_per_sec", "value": v, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["workload"] + " - CPU-port sample; per-sample rate extrapolates to " + cfg["workload"],
                       "sample_of": cfg["workload"], "T": c["T"], "N": c["N"], "B_ppo": c["B_ppo"], "B_gail": c["B_gail"],
                       "ppo_epoch": c["ppo_epoch"], "gail_epoch": c["gail_epoch"]},
            "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "one full update on " + c["workload"] + " (torch CPU fp32, all host threads) per step; "
                                       "env-steps/s is per-sample work, see cpu_baseline_b512 for the batch-size dependence"},
            "cpu_baseline_b512": b512,
            "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
LAUNCHES_PER_CALL = {"gc_welford_merge": 2, "gc_small_linear_bwd": 2}


class Profiler:
    """Counts C-ABI calls (= our kernel launches) and, when armed, brackets every tensor-core contraction with CUDA
    events on the launching stream, recording its algorithmic FLOPs."""

    def __init__(self, A):
        self.A, self.calls, self.launches, self.records, self.armed = A, 0, 0, [], False
        self.other = []
        self._orig = A.call
        A.call = self._call

    def _call(self, name, *args):
        self.calls += 1
        self.launches += LAUNCHES_PER_CALL.get(name, 1)
        fl = self._flops(name, args) if self.armed else None
        if fl is None:
            if not self.armed:
                return self._orig(name, *args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = self._orig(name, *args)
            e1.record()
            self.other.append((name, e0, e1))
            return r
        key = name
        if name.startswith("gc_conv"):
            g = args[0]._obj
            key = f"{name}[{g.Cin}->{g.Cout}]"
        elif name.startswith("gc_linear"):
            key = f"{name}[{'x'.join(str(v) for v in (args[7:10] if name == 'gc_linear_fwd' else args[9:12] if name == 'gc_linear_dgrad' else args[6:9]))}]"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = self._orig(name, *args)
        e1.record()
        self.records.append((key, fl, e0, e1, self._bytes(name, args)))
        return r

    @staticmethod
    def _bytes(name, a):
        """Algorithmic HBM bytes of a conv contraction: each operand read once, the result written once (fp32)."""
        if name in ("gc_conv_fprop", "gc_conv_dgrad", "gc_conv_wgrad"):
            g = a[0]._obj
            x = 4.0 * g.B * g.H * g.W * g.Cin
            y = 4.0 * g.B * g.OH * g.OW * g.Cout
            return x + y          # fprop: read x, write y | dgrad: read dy, write dx | wgrad: read dy and x (dw is tiny)
        return 0.0

    @staticmethod
    def _flops(name, a):
        if name == "gc_conv_fprop" or name == "gc_conv_dgrad" or name == "gc_conv_wgrad":
            g = a[0]._obj
            return 2.0 * g.B * g.OH * g.OW * g.Cout * g.KH * g.KW * g.Cin
        if name == "gc_linear_fwd":
            return 2.0 * a[7] * a[8] * a[9]
        if name == "gc_linear_dgrad":
            return 2.0 * a[9] * a[10] * a[11]
        if name == "gc_linear_wgrad":
            return 2.0 * a[6] * a[7] * a[8]
        return None

    def summary(self):
        torch.cuda.synchronize()
        by = {}
        for name, fl, e0, e1, nbytes in self.records:
            t = e0.elapsed_time(e1) * 1e-3
            f, tt, n, bb = by.get(name, (0.0, 0.0, 0, 0.0))
            by[name] = (f + fl, tt + t, n + 1, bb + nbytes)
        tot_f = sum(v[0] for v in by.values()); tot_t = sum(v[1] for v in by.values())
        oth = {}
        for name, e0, e1 in self.other:
            t, n = oth.get(name, (0.0, 0))
            oth[name] = (t + e0.elapsed_time(e1) * 1e-3, n + 1)
        self.other_summary = {k: {"seconds": v[0], "launches": v[1]} for k, v in sorted(oth.items(), key=lambda kv: -kv[1][0])}
        return by, tot_f, tot_t


def hbm_microbench(A, dev, pk):
    """GAE scan and fused PPO-loss kernels alone (BASELINE.json configs[1]: 16 envs x 2048 steps) plus the asymptotic
    size where the launch is long enough for HBM bandwidth to be the bound; CUDA events, L2 flushed between launches."""
    out = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, reps=10):
        ts = []
        for i in range(reps + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e-3)
        return statistics.median(ts)

    for T, N in ((2048, 16), (4096, 18944)):   # configs[1] size, and an asymptotic size (N = 148 SMs x 4 CTAs x 32 envs)
        r = torch.rand(T, N, 1, device=dev); v = torch.randn(T + 1, N, 1, device=dev); m = torch.ones(T + 1, N, 1, device=dev)
        ret = torch.zeros_like(v)
        t = timeit(lambda: A.gae_returns(r, v, m, ret, 0.99, 0.95))
        b = 16.0 * T * N
        out.append(dict(kernel="gae_scan", T=T, N=N, bytes=b, seconds=t, gbs=b / t / 1e9, frac=b / t / 1e9 / pk["hbm"]))
        del r, v, m, ret
    for B in (4096, 1 << 24):
        head = torch.randn(B, 4, device=dev); act = torch.randn(B, 2, device=dev); s = [torch.randn(B, device=dev) for _ in range(3)]
        stats = torch.tensor([0.0, float(B), float(B), 0.0], dtype=torch.float64, device=dev)
        dh = torch.zeros(B, 4, device=dev); acc = torch.zeros(4, dtype=torch.float64, device=dev)
        t = timeit(lambda: A.ppo_loss(head, act, s[0], s[1], s[2], None, stats, dh, None, None, acc, B, HP["logstd"], True, 0.1, 0.5, 1.0, 0))
        b = 48.0 * B + 4.0 * B   # 48 B/sample (log-prob fused) + the pad column of the [B,4] head rows
        out.append(dict(kernel="ppo_loss_fwd_bwd", B=B, bytes=b, seconds=t, gbs=b / t / 1e9, frac=b / t / 1e9 / pk["hbm"]))
        del head, act, s, dh
    return out


def cpu_baseline_sample():
    """Oracle port timed on this box's host cores on a bounded sample (one update of configs[0]); rank 0, N=1 only."""
    a = NS(warmup=0, steps=1, gpus=1, config="c1", no_b512=True)
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        run_reference(a)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    return line["cpu_baseline"]


def multi_gpu_parity(dev, rank, world):
    """Hardware check of the multi-GPU semantics: a small rollout of 2*world envs x 16 steps is
    updated once (Discriminator.update + predict_reward + GAE + PPO.update, global minibatches of 16*world rows) by `world`
    ranks in exact-sharding mode, and rank 0 then replays the same update as ONE process on the concatenated envs
    from the same seeds.  Reports the largest parameter / tuple / returns deviations between the two runs (the sharded
    run sums per-rank partial gradients over NCCL, the replay sums them inside one wgrad launch - TF32 products, fp32
    sums in a different order, and Adam's sign-like first steps turn a flipped ~0 gradient into a 2*lr difference)."""
    import torch.distributed as dist
    import gail_carla_b200 as G
    from gail_carla_b200 import synthetic, optim
    from gail_carla_b200.driver import update_iteration
    T, Nl = 16, 2
    Bglob = 16 * world          # ~16 +- 4 rows of every global minibatch per rank (binomial shares; an empty share is legal too)
    N = Nl * world
    sp, asp = NS(shape=(4,)), NS(shape=(2,))

    def build(n_envs, mb, seed=5):
        torch.manual_seed(seed)
        pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False).to(dev)
        agent = G.PPO(pol, HP["clip_param"], 1, mb, HP["value_loss_coef"], dev, lr=HP["lr"], eps=HP["eps"], betas=HP["betas"],
                      max_grad_norm=HP["max_grad_norm"], gamma=None, decay=None, act_space=asp)
        disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, dev, HP["gail_lr"], HP["gail_eps"], HP["gail_betas"],
                               HP["gail_max_grad_norm"]).to(dev)
        return pol, agent, disc

    full = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device=dev, obs_dtype=torch.uint8)
    synthetic.fill_rollout(full, seed=17)
    loader = synthetic.SyntheticExpertLoader(2, Bglob, seed=23, obs_u8=True)      # the same global batches on every rank
    keys = ("obs", "metrics", "actions", "action_log_probs", "value_preds", "returns", "masks", "gail_rewards", "rewards")

    pol, agent, disc = build(Nl, Bglob // world)
    agent.exact_sharding = disc.exact_sharding = True
    ro = G.RolloutStorage(T, Nl, synthetic.OBS_SHAPE, (4,), (2,), device=dev, obs_dtype=torch.uint8)
    for k in keys:
        getattr(ro, k).copy_(getattr(full, k)[:, rank * Nl:(rank + 1) * Nl])
    ro.set_shard(rank, world)
    torch.manual_seed(99)
    d_out, p_out = update_iteration(pol, agent, disc, ro, loader, gamma=HP["gamma"], gae_lambda=HP["gae_lambda"], gail_epoch=1)
    torch.cuda.synchronize()
    dist.barrier()
    res = None
    if rank == 0:
        with optim.single_process():
            pol1, agent1, disc1 = build(N, Bglob)
            torch.manual_seed(99)
            d1, p1 = update_iteration(pol1, agent1, disc1, full, loader, gamma=HP["gamma"], gae_lambda=HP["gae_lambda"], gail_epoch=1)
        torch.cuda.synchronize()

        def pdiff(a, b):
            mx = mean = 0.0
            for (k, x), (_, y) in zip(a.state_dict().items(), b.state_dict().items()):
                d = (x.float() - y.float()).abs()
                mx = max(mx, float(d.max())); mean = max(mean, float(d.mean()))
            return mx, mean
        def tdiff(a, b):
            a = [float("nan") if v is None else float(v) for v in a]; b = [float("nan") if v is None else float(v) for v in b]
            return max((abs(x - y) / (1e-6 + abs(y)) for x, y in zip(a, b) if x == x and y == y), default=0.0)
        pm, pa = pdiff(pol, pol1); dm, da = pdiff(disc, disc1)
        ret = float((ro.returns - full.returns[:, :Nl]).abs().max())
        res = {"config": f"{N} envs x {T} steps over {world} ranks (exact sharding), global minibatch {Bglob}, 2 critic + {T * N // Bglob} PPO steps, "
                         "vs a one-process replay of the concatenated envs on rank 0",
               "policy_params_max_abs_diff": pm, "policy_params_max_abs_diff_over_lr": pm / HP["lr"],
               "policy_params_worst_tensor_mean_abs_diff_over_lr": pa / HP["lr"],
               "critic_params_max_abs_diff": dm, "critic_params_max_abs_diff_over_lr": dm / HP["gail_lr"],
               "critic_params_worst_tensor_mean_abs_diff_over_lr": da / HP["gail_lr"],
               "disc_update_tuple_max_rel_diff": tdiff(d_out[0], d1[0]), "ppo_update_tuple_max_rel_diff": tdiff(p_out, p1),
               "returns_max_abs_diff_rank0_shard": ret}
    dist.barrier()
    del full, ro, pol, agent, disc
    from gail_carla_b200 import engine as E
    E.release_workspaces()
    return res


def run_b200(args):
    import torch.distributed as dist
    import gail_carla_b200 as G
    from gail_carla_b200 import _abi as A, synthetic
    from gail_carla_b200.driver import update_iteration

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; stdout must carry only the JSON
        # line, so file descriptor 1 points at stderr while the process group and its first collective are set up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            import datetime
            # a collective that does not complete within 5 minutes is a bug: fail the run instead of stalling it
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=5))
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    A.load_library()
    if args.no_graphs:
        from gail_carla_b200 import graphs as _graphs
        _graphs.ENABLED = False
    parity_multi = None
    if world > 1 and not args.no_parity:
        try:
            parity_multi = multi_gpu_parity(dev, rank, world)
        except Exception as ex:      # a failed check is reported in the line, it must not take the measurement down
            parity_multi = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
    prof = Profiler(A)
    pk = peaks()
    c = dict(CONFIGS[args.config])
    T, N_total = c["T"], c["N"]
    if N_total % world or c["B_ppo"] % world or c["B_gail"] % world:
        raise SystemExit(f"config {args.config} does not split over {world} ranks")
    N = N_total // world                      # envs are sharded across ranks
    Bp, Bg = c["B_ppo"] // world, c["B_gail"] // world
    sp, asp = NS(shape=(4,)), NS(shape=(2,))
    torch.manual_seed(1)                      # identical replicas on every rank
    pol = G.Policy(synthetic.OBS_SHAPE, sp, asp, True, HP["logstd"], False).to(dev)
    agent = G.PPO(pol, HP["clip_param"], c["ppo_epoch"], Bp, HP["value_loss_coef"], dev, lr=HP["lr"], eps=HP["eps"],
                  betas=HP["betas"], max_grad_norm=HP["max_grad_norm"], gamma=None, decay=None, act_space=asp)
    disc = G.Discriminator(synthetic.OBS_SHAPE, sp, asp, 100, dev, HP["gail_lr"], HP["gail_eps"], HP["gail_betas"],
                           HP["gail_max_grad_norm"]).to(dev)
    u8 = args.obs_store == "u8"
    # Observations are uint8/255 by construction (carla_env.py:134-138): the byte store is lossless.  --obs-store f32
    # keeps the reference's fp32 layout (4x the HBM, PCIe and gather traffic).
    ro = G.RolloutStorage(T, N, synthetic.OBS_SHAPE, (4,), (2,), device=dev, obs_dtype=torch.uint8 if u8 else torch.float32)
    synthetic.fill_rollout(ro, seed=11 + rank, chunk=max(1, 4096 // N))
    n_batches = (T * N) // Bg
    loader = synthetic.SyntheticExpertLoader(n_batches, Bg, seed=21 + rank, pin=True, obs_u8=u8)
    torch.manual_seed(100 + rank)             # minibatch permutations / mix-up alphas differ per shard

    def step():
        return update_iteration(pol, agent, disc, ro, loader, gamma=HP["gamma"], gae_lambda=HP["gae_lambda"],
                                gail_epoch=c["gail_epoch"], bcgail=False, diagnostics=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = A.LAUNCHES          # C-ABI kernel launches, graph replays included (gail_carla_b200/graphs.py)
    dt = timed(step, args.steps)
    launches = A.LAUNCHES - launches0
    clocks = sampler.stop() if sampler else None
    env_steps = T * N_total
    value = env_steps * args.steps / dt

    # ---- end-to-end: the rollout lives in pinned HOST memory (as tools/storage.py keeps it) and is uploaded every step
    names = ("obs", "metrics", "actions", "action_log_probs", "value_preds", "masks")
    host = {}
    try:
        for k in names:
            src = ro.flat(k).view(getattr(ro, k).shape)          # plain tensor view (ByteObs -> uint8)
            host[k] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            host[k].copy_(src)
        pinned = True
    except RuntimeError:
        host = {k: ro.flat(k).view(getattr(ro, k).shape).cpu() for k in names}
        pinned = False
    h2d = sum(v.numel() * v.element_size() for v in host.values()) + sum(t.numel() * t.element_size() for b in loader for t in b)
    result_host = torch.empty(T, N, 1, dtype=torch.float32, pin_memory=True)

    # Double-buffered upload: rollout i+1 is copied from pinned host memory into a second obs buffer while update i
    # runs.  Host->device transfers are served in submission order by one DMA queue (55 GB/s, idle or under load:
    # profiles/r01_h2d_bandwidth.txt) that the expert-batch prefetch of Discriminator.update also uses, so the rollout
    # goes up in chunks: one chunk is enqueued (on its own stream) each time the discriminator pulls the next expert batch,
    # sized so that chunk + expert batch fit inside one minibatch of compute, and the rest when the discriminator epochs
    # are over, overlapping predict_reward + GAE + PPO.update.  This is how the storage is fed in the reference flow too
    # - insert() copies one time slice per env step while the simulator runs (tools/learn.py:111-133).  Every timed step
    # issues and completes one full rollout upload; the small tensors go on the compute stream.
    up_stream = torch.cuda.Stream(device=dev)
    obs_plain = ro.flat("obs").view(ro.obs.shape)
    try:
        obs_bufs = [obs_plain, torch.empty_like(obs_plain)]
        double = True
    except RuntimeError:                                   # no room for a second obs buffer: upload, then update
        obs_bufs = [obs_plain, obs_plain]
        double = False
    small = [k for k in names if k != "obs"]
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "chunk": 0}
    flat_host = host["obs"].view(-1)
    NCH = 24
    per = (flat_host.numel() + NCH - 1) // NCH

    def enqueue_chunks(j, upto):
        """Copy chunks [state.chunk, upto) of the next rollout into obs_bufs[j]; record the event after the last one."""
        flat = obs_bufs[j].view(-1)
        with torch.cuda.stream(up_stream):
            while state["chunk"] < min(upto, NCH):
                o = state["chunk"] * per
                flat[o:o + per].copy_(flat_host[o:o + per], non_blocking=True)
                state["chunk"] += 1
            if state["chunk"] == NCH:
                uploaded[j].record(up_stream)

    def enqueue_upload(j):
        up_stream.wait_stream(torch.cuda.current_stream())  # (serial mode: the buffer is the live one)
        state["chunk"] = 0
        enqueue_chunks(j, NCH)

    class InterleavedLoader:
        """The expert loader, with one rollout chunk enqueued ahead of every expert batch after the first."""
        def __init__(self, inner):
            self.inner, self.batch_size = inner, inner.batch_size
        def __len__(self):
            return len(self.inner)
        def __iter__(self):
            for k, batch in enumerate(self.inner):
                # 15 of the 24 chunks ride along with the expert batches, the other 9 go up under rewards + PPO (measured:
                # moving more of them behind the discriminator epochs does not help and slows the expert-resident variant)
                if double and k > 0 and state["chunk"] < NCH - 8:
                    enqueue_chunks((state["i"] + 1) % 2, state["chunk"] + 1)
                yield batch

    rewards_orig = disc.predict_rewards_rollout

    def rewards_hook(rollouts):
        if double:
            enqueue_chunks((state["i"] + 1) % 2, NCH)
        return rewards_orig(rollouts)

    e2e_loader = {"l": InterleavedLoader(loader)}

    def step_inner():
        return update_iteration(pol, agent, disc, ro, e2e_loader["l"], gamma=HP["gamma"], gae_lambda=HP["gae_lambda"],
                                gail_epoch=c["gail_epoch"], bcgail=False, diagnostics=False)

    def step_e2e():
        i = state["i"]
        if not double:
            enqueue_upload(0)
        torch.cuda.current_stream().wait_event(uploaded[i % 2])   # rollout i is resident in obs_bufs[i % 2]
        ro.obs = obs_bufs[i % 2]
        for k in small:
            getattr(ro, k).copy_(host[k], non_blocking=True)
        state["chunk"] = 0
        out = step_inner()                                # tuples are read back inside (one D2H per update call)
        result_host.copy_(ro.returns[:-1], non_blocking=True)   # the step's result tensor back to the host
        torch.cuda.current_stream().synchronize()
        state["i"] = i + 1
        return out

    def timed_e2e(k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            step_e2e()
        torch.cuda.current_stream().wait_stream(up_stream)   # the k-th upload finishes inside the timed region
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    disc.predict_rewards_rollout = rewards_hook
    if double:
        enqueue_upload(0)
    step_e2e()
    dt_e2e = timed_e2e(args.steps)
    torch.cuda.synchronize()
    e2e = env_steps * args.steps / dt_e2e
    h2d_rollout = sum(v.numel() * v.element_size() for v in host.values())

    # ---- same, with the (static) expert data set resident in HBM as the uint8 bytes its PNGs hold (SURVEY 8f row 3,
    # gail_carla_b200/expert.py): only the rollout crosses PCIe.  Reported next to `e2e`, not instead of it.
    e2e_res = None
    try:
        from gail_carla_b200.expert import DeviceExpertLoader, ExpertDataset
        ds = ExpertDataset.from_tensors(torch.cat([b[0] if b[0].dtype == torch.uint8 else (b[0] * 255.0).round().to(torch.uint8) for b in loader]),
                                        torch.cat([b[1] for b in loader]), torch.cat([b[2] for b in loader]))
        e2e_loader["l"] = InterleavedLoader(DeviceExpertLoader(ds, Bg, shuffle=False, drop_last=True, device=dev))
        del ds
        step_e2e()
        dt_res = timed_e2e(args.steps)
        torch.cuda.synchronize()
        e2e_res = {"value": env_steps * args.steps / dt_res, "unit": "env-steps/s", "ms_per_step": dt_res / args.steps * 1e3,
                   "h2d_bytes_per_step": h2d_rollout * world, "d2h_bytes_per_step": (result_host.numel() * 4 + 8 * 15) * world,
                   "note": "expert data set resident in HBM as uint8 (uploaded once, outside the timed region); the rollout "
                           "is uploaded from pinned host memory every step as in `e2e`"}
    except RuntimeError as ex:                              # e.g. no room for the table
        e2e_res = {"unavailable": str(ex)[:200]}
    e2e_loader["l"] = None
    disc.predict_rewards_rollout = rewards_orig
    ro.obs = obs_bufs[0]
    del obs_bufs[1:]
    d2h = result_host.numel() * 4 + 8 * 15

    # ---- instrumented step: per-contraction CUDA events -> tensor-pipe roofline of the dominant kernel
    from gail_carla_b200 import graphs
    graphs_were = graphs.ENABLED
    graphs.ENABLED = False            # per-kernel events need eager launches
    prof.armed = True
    step()
    prof.armed = False
    graphs.ENABLED = graphs_were
    by, tot_f, tot_t = prof.summary()
    step_s = dt / args.steps
    tf32_peak = pk["bf16"] / 2.0
    # DRAM traffic of the contraction that takes the most time in the step (conv1 fprop, one launch at B=4096), from the
    # committed `ncu --set full` capture (profiles/r01_ncu_traffic.json; algorithmic bytes of that launch: 7.15e9)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp)).get("r01_conv1_fprop_full")
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_src = "profiles/r01_ncu_traffic.json: gc_conv_fprop[16->32] B=4096, one launch (algorithmic 7.15e9 B, %.3f ms)" % (tj["seconds"] * 1e3)
    roof = {"bound": "tensor", "kernel": "umma_gemm_kernel (tcgen05.mma.kind::tf32, all conv/linear contractions)",
            "achieved": tot_f / tot_t / 1e12 if tot_t else None, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": (tot_f / tot_t / 1e12) / tf32_peak if tot_t else None, "traffic": traffic, "traffic_note": traffic_src,
            "peak_note": f"TF32 dense = 1/2 of the {pk['src']} sustained bf16 peak ({pk['bf16']} TF/s); MEASURED_PEAKS.json has no TF32 entry",
            "share_of_step": tot_t / step_s if step_s else None, "launches": len(prof.records),
            # mixed roofline: every contraction is bounded by the slower of its tensor time (FLOPs / TF32 peak) and its HBM
            # time (algorithmic bytes / measured HBM peak; conv layers only - the small-channel layers are HBM-bound);
            # mixed_frac = sum of those bounds / sum of the measured times
            "mixed_ideal_seconds": sum(max(v[0] / (tf32_peak * 1e12), v[3] / (pk["hbm"] * 1e9)) for v in by.values()),
            "mixed_frac": (sum(max(v[0] / (tf32_peak * 1e12), v[3] / (pk["hbm"] * 1e9)) for v in by.values()) / tot_t) if tot_t else None,
            # per contraction: tensor-pipe rate, and for the convolutions the algorithmic HBM rate (the small-channel layers
            # conv1 / conv2 are HBM-bound: ~17 and ~80 FLOP per byte moved)
            "per_op": {k: dict({"tflops": v[0] / v[1] / 1e12, "seconds": v[1], "launches": v[2]},
                               **({"hbm_gbs": v[3] / v[1] / 1e9, "hbm_frac": v[3] / v[1] / 1e9 / pk["hbm"]} if v[3] else {}))
                       for k, v in by.items()}}

    hbm = hbm_microbench(A, dev, pk) if (rank == 0 and world == 1) else None   # before the power-hungry GEMM yardstick
    if rank == 0 and world == 1:
        # yardstick only (not on any product path): what cuBLAS sustains with TF32 inputs on this box, since
        # MEASURED_PEAKS.json has no TF32 entry and 1/2 x (sustained bf16) is an assumption
        try:
            old = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = True
            a_ = torch.randn(8192, 8192, device=dev); b_ = torch.randn(8192, 8192, device=dev)
            for _ in range(3):
                torch.matmul(a_, b_)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                torch.matmul(a_, b_)
            e1.record(); torch.cuda.synchronize()
            roof["tf32_cublas_sustained_tflops"] = 40 * 2.0 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
            torch.backends.cuda.matmul.allow_tf32 = old
            del a_, b_
        except Exception as ex:   # pragma: no cover
            roof["tf32_cublas_sustained_tflops"] = None
    obs_gb = ro.obs.numel() * ro.obs.element_size() / 1e9
    gpu_base = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        # yardstick, measured last with every buffer of the product path released: stock PyTorch on this same GPU
        del ro, loader, host, obs_bufs
        from gail_carla_b200 import engine as E
        E.release_workspaces()
        try:
            gpu_base = gpu_baseline_sample(dev)
            if gpu_base and "value" in gpu_base:
                gpu_base["speedup_of_value"] = value / gpu_base["value"]
        except Exception as ex:   # pragma: no cover - the yardstick must never take the bench line down
            gpu_base = {"unavailable": f"{type(ex).__name__}: {str(ex)[:160]}"}
    if rank == 0:
        cpu = cpu_baseline_sample() if world == 1 and not args.no_cpu_baseline else None
        line = {"metric": "ppo_wdgail_update_env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
                "config": {"workload": c["workload"], "T": T, "N": N_total, "B_ppo": c["B_ppo"], "B_gail": c["B_gail"],
                           "ppo_epoch": c["ppo_epoch"], "gail_epoch": c["gail_epoch"], "envs_per_rank": N,
                           "obs_store": "uint8 bytes (lossless: observations are uint8/255 by construction)" if u8 else "fp32 (reference layout)",
                           "l2": "inputs (rollout obs, %.1f GB per rank) are larger than L2" % obs_gb,
                           "cuda_graphs": not args.no_graphs,
                           "parallelism": f"env-sharded data parallel x{world}, NCCL grad all-reduce (two buckets, overlapped with the conv backward)"},
                "e2e": {"value": e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                        "ms_per_step": dt_e2e / args.steps * 1e3, "pinned": pinned,
                        "upload": ("double-buffered" if double else "serial") + ": rollout i+1 is copied from pinned host memory behind "
                                  "predict_reward + PPO.update of step i; one full rollout upload + all expert batches per timed step"},
                "e2e_expert_resident": e2e_res,
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_hbm_kernels": hbm, "cpu_baseline": cpu,
                "gpu_baseline": gpu_base, "parity_multi_gpu": parity_multi,
                "other_kernels_seconds": prof.other_summary}
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL resources: release them before the communicator goes away (see graphs.release_all);
        # the line is out and every rank is done, so leave without the interpreter's teardown - a hang there (seen once with
        # live graphs: the NCCL watchdog blocked in an event destroy for minutes) must never cost the run
        from gail_carla_b200 import graphs as _g
        torch.cuda.synchronize()
        dist.barrier()
        _g.release_all()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-b512", action="store_true", help="reference arm: skip the one-off B=512 CPU sample")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-PyTorch-on-CUDA yardstick")
    ap.add_argument("--obs-store", default="u8", choices=["u8", "f32"],
                    help="rollout observation store: uint8 bytes (lossless, observations are uint8/255) or the reference's fp32")
    ap.add_argument("--no-parity", action="store_true", help="multi-GPU runs: skip the exact-sharding parity check")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
