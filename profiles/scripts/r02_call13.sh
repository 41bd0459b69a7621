#!/bin/bash
mkdir -p gpurun_out
GC_UMMA_STATS=1 B=4096 REPS=2 timeout 90 python tests/gpu_probe_layers.py > gpurun_out/r02_sw_stats.log 2>&1
echo done
