#!/bin/bash
# Round-2 GPU call 16: split accumulation chains (default build) and the one-op-per-link D chain (variant) on truncated c2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu16.log
: > $L
timeout 300 python tools/gpu_perf_c2.py 640 48 base 2>&1 | grep -E "^\[|vs base|rror" >> $L
PERF_LIB=tools/bin/libwitch_chain1.so timeout 300 python tools/gpu_perf_c2.py 640 48 chain1 2>&1 | grep -E "^\[|vs base|rror" >> $L
timeout 300 python tools/gpu_perf_c2.py 640 48 base2 2>&1 | grep -E "^\[|vs base|rror" >> $L
rm -f gpurun_out/scores_*.npz
cat $L
