#!/bin/bash
# round 2, call 10: GPU tests after the loop goldens / gradient re-attachment changes (no kernel change since call 8)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest_gpu_final3.log
echo done
