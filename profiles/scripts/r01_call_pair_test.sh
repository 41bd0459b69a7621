mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/t_pair.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_pair.log
for L in 3 4; do for OP in fprop dgrad; do GC_UMMA_STATS=1 B=4096 REPS=1 LAYER=$L OP=$OP timeout 120 python tests/gpu_probe_one.py 2>&1 | grep "umma-stats" | tail -1 | cut -c1-330; done; done
for L in 3 4; do for OP in fprop dgrad; do
 for NO in 0 1; do if [ $NO = 1 ]; then export GC_NO_CTA2=1; else unset GC_NO_CTA2; fi
 B=4096 REPS=5 LAYER=$L OP=$OP timeout 120 python - <<'P'
import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from gail_carla_b200 import _abi as A, engine as E
B = 4096; layer = int(os.environ["LAYER"]); op = os.environ["OP"]
g = E.conv_geom(layer, B); cin, cout = E.CONV_CH[layer - 1], E.CONV_CH[layer]
nw = cout * cin * 16
x = torch.randn(B * g.in_batch_stride, device="cuda"); y = torch.randn(B * g.out_batch_stride, device="cuda")
dx = torch.zeros_like(x); w = torch.randn(nw, device="cuda") * 0.05; bias = torch.zeros(cout, device="cuda")
fn = {"fprop": lambda: A.conv_fprop(g, x, w, bias, y, A.EPI_BIAS_LRELU, 0.2), "dgrad": lambda: A.conv_dgrad(g, y, w, dx, None, 0.2)}[op]
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 5
fl = 2.0 * B * g.OH * g.OW * cout * g.KH * g.KW * g.Cin
print(f"conv{layer} {op} pair={'off' if os.environ.get('GC_NO_CTA2') else 'on'}: {t:.3f} ms {fl/t/1e9:.0f} TF/s", flush=True)
P
 done; done; done
