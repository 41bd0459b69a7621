#!/bin/bash
# round 2, call 4 (2 GPUs): GPU tests, conv layer probe with the fused LeakyReLU+bits epilogue, 2-rank bench (NCCL inside the
# captured step, overlapped gradient buckets, multi-GPU parity field)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_pytest_gpu2.log
B=4096 REPS=3 python tests/gpu_probe_layers.py > gpurun_out/r02_layer_probe.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_n2.json 2> gpurun_out/r02_bench_c4_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-graphs > gpurun_out/r02_bench_c4_n2_nographs.json 2> gpurun_out/r02_bench_c4_n2_nographs.err
echo done
