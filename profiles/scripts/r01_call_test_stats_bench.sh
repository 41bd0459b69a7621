mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log
GC_UMMA_STATS=1 B=4096 REPS=1 timeout 300 python tests/gpu_probe_layers.py > gpurun_out/stats_layers3.log 2>&1
for OP in fprop dgrad wgrad; do GC_UMMA_STATS=1 B=4096 REPS=1 LAYER=4 OP=$OP timeout 120 python tests/gpu_probe_one.py >> gpurun_out/stats_layers3.log 2>&1; done
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err
python -c "
import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e_expert_resident'].get('value'),d['roofline']['achieved'],d['roofline']['frac'],d['roofline'].get('tf32_cublas_sustained_tflops'))"
