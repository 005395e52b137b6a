#!/bin/bash
# Round-2 GPU call 14: generation 6 with the C=13 hybrid class (6 parameter sets in registers, 3 in smem) -- A/B on truncated c2, parity tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu14.log
: > $L
WITCH_PARSER=5 timeout 300 python tools/gpu_perf_c2.py 640 48 base 2>&1 | grep -E "^\[|vs base|rror" >> $L
timeout 300 python tools/gpu_perf_c2.py 640 48 gen6c13 2>&1 | grep -E "^\[|vs base|rror" >> $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_pytest14.log >> $L
rm -f gpurun_out/scores_*.npz
cat $L
