mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log
GC_UMMA_STATS=1 B=4096 REPS=1 timeout 300 python tests/gpu_probe_layers.py > gpurun_out/stats_layers3.log 2>&1
B=4096 REPS=3 timeout 300 python tests/gpu_probe_layers.py 2>&1 | tail -3
for OP in fprop dgrad wgrad; do B=4096 REPS=3 LAYER=4 OP=$OP timeout 120 python tests/gpu_probe_one.py 2>&1 | tail -1; done
