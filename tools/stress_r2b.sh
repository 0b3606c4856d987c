#!/bin/bash
# Reproducibility stress of the final round-2 kernels (64 x 64 tile shape): four processes share one B200 and each repeats
# the converged C3 adjustment (nested dissection) REPS times; tools/stress_repro.py reports any repetition that differs
# from the first beyond FP64 reordering noise.
REPS=${1:-400}
mkdir -p gpurun_out
for t in a b c d; do
  python tools/stress_repro.py C3 $REPS $t nd > gpurun_out/stress_r2b_$t.log 2>&1 &
done
wait
cat gpurun_out/stress_r2b_*.log | grep -E "MISMATCH|repetitions" 
