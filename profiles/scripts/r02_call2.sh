#!/bin/bash
# round 2, call 2: new GPU tests, parity calibration probe, bench with the uint8 store, gather microbench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu.log
python tests/gpu_probe_parity.py > gpurun_out/r02_parity_probe.log 2>&1
python profiles/scripts/hbm_probe.py > gpurun_out/r02_hbm_probe2.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_u8.json 2> gpurun_out/r02_bench_c4_u8.err
echo done
