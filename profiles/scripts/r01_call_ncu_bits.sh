mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
BITS=1 B=4096 LAYER=1 OP=fprop REPS=2 timeout 120 python tests/gpu_probe_one.py > gpurun_out/plain_one.log 2>&1 && \
BITS=1 B=4096 LAYER=1 OP=fprop REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 1 -c 1 -o gpurun_out/r01_conv1_fprop_bits_full python tests/gpu_probe_one.py > gpurun_out/ncu_one_bits.log 2>&1
ls -la gpurun_out/*.ncu-rep
