# launch list of the bench command (c4 batch shapes on a short rollout) + full ncu captures of the top contractions
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 python bench.py --config c4s --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_c4s.log 2>gpurun_out/plain_c4s.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_c4s.csv python bench.py --config c4s --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c4s.log 2>&1
for spec in "1 fprop" "2 fprop" "3 fprop" "2 wgrad" "1 wgrad" "4 dgrad"; do set -- $spec
B=4096 LAYER=$1 OP=$2 REPS=2 timeout 120 python tests/gpu_probe_one.py > gpurun_out/plain_one.log 2>&1 && \
B=4096 LAYER=$1 OP=$2 REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_gemm -s 1 -c 1 -o gpurun_out/r01_conv$1_$2_full python tests/gpu_probe_one.py > gpurun_out/ncu_one_$1_$2.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | head
