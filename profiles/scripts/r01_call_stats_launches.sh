mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
GC_UMMA_STATS=1 B=4096 REPS=1 timeout 300 python tests/gpu_probe_layers.py > gpurun_out/stats_layers.log 2>&1
for L in 4; do for OP in fprop dgrad wgrad; do GC_UMMA_STATS=1 B=4096 REPS=1 LAYER=$L OP=$OP timeout 120 python tests/gpu_probe_one.py >> gpurun_out/stats_layers.log 2>&1; done; done
timeout 600 python bench.py --config c4s --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c4s.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_c4s.csv python bench.py --config c4s --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4s.log 2>&1
B=4096 LAYER=2 OP=fprop REPS=2 timeout 120 python tests/gpu_probe_one.py > gpurun_out/plain_one.log 2>&1 && \
B=4096 LAYER=2 OP=fprop REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 1 -c 1 -o gpurun_out/r01_conv2_fprop_full python tests/gpu_probe_one.py > gpurun_out/ncu_one.log 2>&1
tail -40 gpurun_out/stats_layers.log
