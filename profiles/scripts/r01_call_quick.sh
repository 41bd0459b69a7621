mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_quick.log
B=4096 REPS=3 timeout 300 python tests/gpu_probe_layers.py 2>&1 | tail -3
for OP in fprop dgrad wgrad; do B=4096 REPS=5 LAYER=4 OP=$OP timeout 120 python - <<'P'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from gail_carla_b200 import _abi as A, engine as E
B = 4096; layer = 4; op = os.environ["OP"]
g = E.conv_geom(layer, B); cin, cout = E.CONV_CH[layer - 1], E.CONV_CH[layer]
nw = cout * cin * 16
x = torch.randn(B * g.in_batch_stride, device="cuda"); y = torch.randn(B * g.out_batch_stride, device="cuda")
dx = torch.zeros_like(x); w = torch.randn(nw, device="cuda") * 0.05; bias = torch.zeros(cout, device="cuda")
z = A.conv_wgrad_splits(g); part = torch.zeros(z * nw, device="cuda")
fn = {"fprop": lambda: A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2), "dgrad": lambda: A.conv_dgrad(g, y, w, dx, None, 0.2),
      "wgrad": lambda: A.conv_wgrad(g, y, x, part, z)}[op]
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5
fl = 2.0 * B * g.OH * g.OW * cout * g.KH * g.KW * g.Cin
print(f"conv4 {op}: {t:.3f} ms {fl/t/1e9:.0f} TF/s", flush=True)
P
done
