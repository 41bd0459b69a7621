#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_bench_c4_final3.json 2> gpurun_out/r02_bench_c4_final3.err
echo done
