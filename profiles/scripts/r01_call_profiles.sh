# launch list of the bench command (c4 batch shapes on a short rollout) + full ncu captures of the top contractions
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --config c4s --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c4s.log 2>gpurun_out/plain_c4s.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_c4s.csv python bench.py --config c4s --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4s.log 2>&1
for spec in "1 fprop" "2 fprop" "3 fprop" "2 wgrad" "1 wgrad"; do set -- $spec
B=4096 LAYER=$1 OP=$2 REPS=2 timeout 120 python tests/gpu_probe_one.py > gpurun_out/plain_one.log 2>&1 && \
B=4096 LAYER=$1 OP=$2 REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 1 -c 1 -o gpurun_out/r01_conv$1_$2_full python tests/gpu_probe_one.py > gpurun_out/ncu_one_$1_$2.log 2>&1
done
python - <<'P'
import torch, time
# H2D bandwidth from pinned memory: idle, and while a long GEMM stream keeps the SMs busy
h = torch.empty(1 << 30, dtype=torch.float32, pin_memory=True); d = torch.empty_like(h, device="cuda")
def bw():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        e0.record(s); d.copy_(h, non_blocking=True); e1.record(s)
    e1.synchronize(); return 4.295 / (e0.elapsed_time(e1) * 1e-3)
print("H2D idle GB/s", bw(), bw())
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
x = torch.randn(1 << 28, device="cuda"); y = torch.empty_like(x)
for _ in range(40): torch.mm(a, b)
print("H2D under fp32 GEMM GB/s", bw())
torch.cuda.synchronize()
for _ in range(300): y.copy_(x)
print("H2D under HBM copy stream GB/s", bw())
torch.cuda.synchronize()
P
ls -la gpurun_out/*.ncu-rep | head
