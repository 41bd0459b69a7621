#!/bin/bash
# round 2, call 1: baseline of the round-1 binary (3-set prefetch ring e2e) + ncu --set full of every HBM-bound kernel
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_before.json 2> gpurun_out/r02_bench_c4_before.err
python profiles/scripts/hbm_probe.py > gpurun_out/r02_hbm_probe_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'gae_scan|ppo_loss|clip_adam|grad_sumsq|gather_obs|colsum|mixup|grad_penalty' -c 60 \
    -o gpurun_out/r02_hbm_kernels python profiles/scripts/hbm_probe.py > gpurun_out/r02_hbm_probe_ncu.log 2>&1
echo done
