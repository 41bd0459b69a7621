#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_learn_gpu.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_pytest_learn_gpu.log
echo done
