#!/usr/bin/env python
"""Launch every HBM-bound kernel of the hot path twice at its named size and at an asymptotic size.

Run plain first, then under `ncu --set full` (see profiles/README.md); `summarize_ncu.py` turns the report into
achieved GB/s (algorithmic bytes / gpu__time_duration) and DRAM traffic per launch.  Sizes and algorithmic bytes are
printed as JSON lines so the summary can be joined with the ncu launch order.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gail_carla_b200 import _abi as A  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
LOG = []


def note(kernel, label, nbytes, launches=2):
    LOG.append(dict(kernel=kernel, label=label, algorithmic_bytes=nbytes, launches=launches))


class timed:
    """CUDA-event time of the enclosed launches (the second, warm one of each pair is what the JSON line reports)."""
    last_ms = None

    def __enter__(self):
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.e0.record()
        return self

    def __exit__(self, *a):
        self.e1.record(); torch.cuda.synchronize()
        timed.last_ms = self.e0.elapsed_time(self.e1)


def gae(T, N, label):
    r = torch.rand(T, N, 1, device=dev); v = torch.randn(T + 1, N, 1, device=dev); m = (torch.rand(T + 1, N, 1, device=dev) > 0.02).float()
    ret = torch.zeros_like(v)
    for _ in range(2):
        A.gae_returns(r, v, m, ret, 0.99, 0.95)
    note("gae_scan_kernel", label, 16.0 * T * N)


def ppo(B, label):
    head = torch.randn(B, 4, device=dev); act = torch.randn(B, 2, device=dev); s = [torch.randn(B, device=dev) for _ in range(3)]
    stats = torch.tensor([0.0, float(B), float(B), 0.0], dtype=torch.float64, device=dev)
    dh = torch.zeros(B, 4, device=dev); acc = torch.zeros(4, dtype=torch.float64, device=dev)
    for _ in range(2):
        A.ppo_loss(head, act, s[0], s[1], s[2], None, stats, dh, None, None, acc, B, [-1.4, -3.2], True, 0.1, 0.5, 1.0, 0)
    note("ppo_loss_kernel", label, 52.0 * B)


def adam(n, label):
    p = torch.randn(n, device=dev); g = torch.randn(n, device=dev) * 1e-3; m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    ss = torch.zeros(1, dtype=torch.float64, device=dev)
    for _ in range(2):
        ss.zero_()
        A.grad_sumsq(g, n, ss)
        A.clip_adam(p, g, m, v, n, ss, 0.5, 1e-4, 0.9, 0.99, 1e-8, 0.1, 0.01)
    note("grad_sumsq_kernel", label, 4.0 * n)
    note("clip_adam_kernel", label, 28.0 * n)


def gather(B, rows, label, u8):
    if u8:
        src = torch.randint(0, 256, (rows, 3, 192, 192), dtype=torch.uint8, device=dev)
    else:
        src = torch.rand(rows, 3, 192, 192, device=dev)
    idx = torch.randperm(rows, device=dev)[:B].contiguous()
    out = torch.empty(B, A.S2D_PER_SAMPLE, device=dev)
    for _ in range(2):
        with timed():
            A.gather_obs_s2d(src, idx, out, B)
    per = 3 * 192 * 192 * (1 if u8 else 4) + 4 * A.S2D_PER_SAMPLE
    note("gather_obs_u8_s2d_kernel" if u8 else "gather_obs_s2d_kernel", label, float(per) * B)
    LOG[-1]["event_ms"] = timed.last_ms; LOG[-1]["event_gbs"] = float(per) * B / timed.last_ms / 1e6


def colsum(rows, Cc, label):
    x = torch.randn(rows, Cc, device=dev); out = torch.zeros(Cc, device=dev)
    for _ in range(2):
        A.colsum(x, Cc, rows, Cc, out)
    note("colsum_kernel", label, 4.0 * rows * Cc)


def mix_and_penalty(B, label):
    per = A.S2D_PER_SAMPLE
    xe = torch.randn(B, per, device=dev); xp = torch.randn(B, per, device=dev); out = torch.empty(B, per, device=dev)
    alpha = torch.rand(B, device=dev)
    for _ in range(2):
        A.mixup(xe, xp, alpha, out, B, per)
    note("mixup_kernel", label, 12.0 * B * per)
    acc = torch.zeros(4, dtype=torch.float64, device=dev)
    for _ in range(2):
        A.grad_penalty(xe, out, acc, B, per, 10.0, (1 / 0.229, 1 / 0.224, 1 / 0.225))
    note("grad_penalty_kernel", label, 12.0 * B * per)


def main():
    A.load_library()
    gae(2048, 16, "configs[1]: 16 envs x 2048 steps")
    gae(1024, 64, "configs[3]: 64 envs x 1024 steps")
    gae(4096, 18944, "asymptotic: 18944 envs x 4096 steps")
    ppo(4096, "one minibatch B=4096")
    ppo(1 << 24, "asymptotic B=16.8M")
    adam(14462003 // 64 * 64 + 64 * 12, "policy parameters (14.46 M)")
    gather(4096, 8192, "B=4096 from an 8192-row fp32 table", False)
    gather(4096, 8192, "B=4096 from an 8192-row uint8 table", True)
    colsum(4096 * 46 * 46, 64, "delta_2 of a B=4096 minibatch")
    mix_and_penalty(2048, "B=2048 s2d images")
    torch.cuda.synchronize()
    for rec in LOG:
        print(json.dumps(rec))


if __name__ == "__main__":
    main()
