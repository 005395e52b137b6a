#!/bin/bash
# Round-2 GPU call 10: parser generations 5 (default), 6 (C=16, 4-warp CTAs, parameters in smem), 7 (C=12, 5-warp CTAs) on truncated c2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu10.log
: > $L
WITCH_PARSER=5 timeout 300 python tools/gpu_perf_c2.py 640 48 base 2>&1 | grep -E "^\[|vs base" >> $L
for g in 6 7; do WITCH_PARSER=$g timeout 300 python tools/gpu_perf_c2.py 640 48 gen$g 2>&1 | grep -E "^\[|vs base|rror" >> $L; done
WITCH_PARSER=6 timeout 600 python -m pytest tests -m gpu -q -x -k "scores_weights or properties or device_pipeline or c1_cuda or live" > gpurun_out/r2_pytest10_gen6.log 2>&1; echo "pytest gen6 rc=$?" >> $L
tail -3 gpurun_out/r2_pytest10_gen6.log >> $L
rm -f gpurun_out/scores_*.npz
cat $L
