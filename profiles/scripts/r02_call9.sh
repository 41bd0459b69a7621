#!/bin/bash
# round 2, call 9 (2 GPUs): the final binary through the multi-GPU path (graphs + NCCL buckets + parity check with 16 rows per rank)
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_n2_final.json 2> gpurun_out/r02_bench_c4_n2_final.err
echo "rc=$?" >> gpurun_out/r02_bench_c4_n2_final.err
echo done
