#!/bin/bash
# round 2, call 5: GPU tests of the fused column sums / pair gather / split-K un-permute, bench c4 with them, c5 at N=1
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu3.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-gpu-baseline > gpurun_out/r02_bench_c4_fused.json 2> gpurun_out/r02_bench_c4_fused.err
timeout 200 python bench.py --config c4r8 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_bench_c4r8_fused.json 2> gpurun_out/r02_bench_c4r8_fused.err
timeout 600 python bench.py --config c5 --steps 2 --warmup 1 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err
echo done
