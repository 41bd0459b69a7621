#!/bin/bash
# round 2, call 6 (8 GPUs): configs[3] at 8 and 4 ranks, configs[4] at 8 ranks; every command bounded by its own timeout
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_n8.json 2> gpurun_out/r02_bench_c4_n8.err
timeout 240 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_n4.json 2> gpurun_out/r02_bench_c4_n4.err
timeout 300 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --config c5 --steps 2 --warmup 1 > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err
echo done
