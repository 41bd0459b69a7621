#!/bin/bash
# round 2, call 7: final binary - GPU tests, smoke, default bench line, launch list of the c4s profile config
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_final.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/r02_smoke.log 2>&1
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_final.json 2> gpurun_out/r02_bench_c4_final.err
timeout 150 python bench.py --config c4s --steps 1 --warmup 1 --no-graphs --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_plain_c4s.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_c4s.csv \
    python bench.py --config c4s --steps 1 --warmup 1 --no-graphs --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_ncu_c4s.log 2>&1
echo done
