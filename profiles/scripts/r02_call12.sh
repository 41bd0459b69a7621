#!/bin/bash
# round 2, call 12: store-warp epilogue experiment - exactness tests of every contraction path, then the isolated layer timings
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_umma_gpu.py tests/test_fullsize_gpu.py tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_sw_tests.log
B=4096 REPS=3 timeout 90 python tests/gpu_probe_layers.py > gpurun_out/r02_sw_layer_probe.log 2>&1
echo done
