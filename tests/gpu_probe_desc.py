"""Bring-up experiment: properties of the UMMA K-major SWIZZLE_128B shared-memory descriptor that decide whether an
input patch loaded once can be re-used by several convolution taps.
mode 1: descriptor starts r rows (r*128 B) into the tile; mode 2: 8-row groups 1152 B (9 rows) apart."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(mode, r, bo):
    import torch
    from gail_carla_b200 import _abi as A
    g = torch.Generator().manual_seed(1)
    M, N, K = 384, 64, 64
    x = torch.randint(-3, 4, (M, K), generator=g).float(); w = torch.randint(-2, 3, (N, K), generator=g).float()
    y = torch.zeros(M, N, device="cuda")
    A.linear_fwd(x.cuda(), K, w.cuda(), K, None, y, N, M, N, K, 0, 0.2, 1)
    torch.cuda.synchronize()
    full = x @ w.t()
    if mode == 1:
        exp = full[r:r + 128]
    else:
        rows = [9 * gidx + r + i for gidx in range(16) for i in range(8)]
        exp = full[rows]
    got = y[:128].cpu()
    ok = torch.equal(got, exp)
    bad_rows = (got != exp).any(1).nonzero().view(-1).tolist()
    print(f"mode={mode} r={r} base_offset={bo}: {'MATCH' if ok else 'MISMATCH'} bad_rows={bad_rows[:24]} n_bad={len(bad_rows)}", flush=True)

def run_mn(mode, r, bo):
    """MN-major B views: dw[m,n] = sum_k dy[k,m] * x[k+shift(n),n]."""
    import torch
    from gail_carla_b200 import _abi as A
    g = torch.Generator().manual_seed(2)
    M, N, K = 128, 64, 96
    dy = torch.randint(-3, 4, (K, M), generator=g).float(); x = torch.randint(-2, 3, (K, N), generator=g).float()
    dw = torch.zeros(M, N, device="cuda")
    A.linear_wgrad(dy.cuda(), M, x.cuda(), N, dw, N, M, N, K, 1)
    torch.cuda.synchronize()
    xs = torch.cat([x, torch.zeros(16, N)], 0)
    if mode == 3:
        exp = dy.t() @ xs[r:r + K]
    else:   # columns 32..63 are columns 0..31 shifted by one row
        exp = torch.cat([dy.t() @ xs[0:K, :32], dy.t() @ xs[1:K + 1, :32]], 1)
    got = dw.cpu()
    bad = (got != exp)
    print(f"MN-major mode={mode} r={r}: {'MATCH' if not bad.any() else 'MISMATCH'} bad_cols={bad.any(0).nonzero().view(-1).tolist()[:16]} n_bad={int(bad.sum())}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 4 and int(sys.argv[1]) >= 3:
        run_mn(*(int(v) for v in sys.argv[1:]))
    elif len(sys.argv) == 4:
        run(*(int(v) for v in sys.argv[1:]))
    else:
        for mode, r, bo in [(1, 0, 0), (1, 1, 0), (1, 3, 0), (2, 1, 0), (3, 0, 0), (3, 1, 0), (3, 2, 0), (3, 4, 0), (3, 5, 0), (4, 0, 0)]:
            env = dict(os.environ, GC_EXP=f"{mode},{r},{bo}")
            p = subprocess.run([sys.executable, os.path.abspath(__file__), str(mode), str(r), str(bo)], env=env, capture_output=True, text=True, timeout=120)
            print(p.stdout.strip() or ("ERR " + p.stderr[-300:]), flush=True)
