mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_umma_gpu.py tests/test_fullsize_gpu.py tests/test_update_gpu.py tests/test_kernels_gpu.py -x -q > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_quick.log
B=4096 REPS=3 timeout 300 python tests/gpu_probe_layers.py 2>&1 | head -1
GC_NO_SUBTILES=1 B=4096 REPS=3 timeout 300 python tests/gpu_probe_layers.py 2>&1 | head -1
