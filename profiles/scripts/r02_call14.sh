#!/bin/bash
# round 2, call 14: ncu --set full with source-level stall sampling of conv1 fprop + bits (the step's most expensive kernel)
mkdir -p gpurun_out
timeout 60 python tests/gpu_probe_conv1.py > gpurun_out/r02_conv1_plain.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 2 -c 1 -o gpurun_out/r02_conv1_fprop_bits_lean python tests/gpu_probe_conv1.py > gpurun_out/r02_conv1_ncu.log 2>&1
echo done
