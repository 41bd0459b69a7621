#!/bin/bash
# round 2, call 8: re-validation after removing the torch fill kernels from the minibatch path
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_final2.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_bench_c4_final2.json 2> gpurun_out/r02_bench_c4_final2.err
echo done
