# how much of a k-iteration is operand traffic?  (results of the GC_EXP_SKIP runs are numerically wrong by design)
for L in 3 4; do for OP in fprop dgrad; do for SK in 0 1 2 3; do
echo "layer $L $OP skip=$SK"; GC_EXP_SKIP=$SK GC_UMMA_STATS=1 B=4096 REPS=1 LAYER=$L OP=$OP timeout 120 python tests/gpu_probe_one.py 2>&1 | grep umma-stats | tail -1 | cut -c1-260
done; done; done
