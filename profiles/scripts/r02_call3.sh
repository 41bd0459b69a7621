#!/bin/bash
# round 2, call 3: gradient calibration (stock TF32 yardstick, linearised networks), bench with CUDA graphs + uint8 store
mkdir -p gpurun_out
python tests/gpu_probe_parity.py --extra > gpurun_out/r02_parity_probe_extra.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_graphs.json 2> gpurun_out/r02_bench_c4_graphs.err
python bench.py --steps 3 --warmup 3 --no-graphs --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_bench_c4_nographs.json 2> gpurun_out/r02_bench_c4_nographs.err
python bench.py --config c4r8 --steps 3 --warmup 3 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_bench_c4r8_graphs.json 2> gpurun_out/r02_bench_c4r8_graphs.err
python bench.py --config c4r8 --steps 3 --warmup 3 --no-graphs --no-gpu-baseline --no-cpu-baseline > gpurun_out/r02_bench_c4r8_nographs.json 2> gpurun_out/r02_bench_c4r8_nographs.err
echo done
