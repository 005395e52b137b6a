#!/bin/bash
# Round-2 GPU call 3: parity tests with the branch-free trace kernel, parser generations, multi-domain timing per shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu3.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -5 gpurun_out/r2_pytest3.log >> $L
python tools/gpu_perf_c2.py 640 48 base >> $L 2>&1
for g in 2 3 4; do WITCH_PARSER=$g timeout 300 python tools/gpu_perf_c2.py 640 48 gen$g >> $L 2>&1 || echo "gen$g FAILED" >> $L; done
for v in pf2 pf12; do
  echo "== md prefetch variant $v (c2 slab)" >> $L
  PERF_LIB=tools/bin/libwitch_$v.so HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 300 python tools/gpu_hosttime.py c2 4 2>&1 | grep -E "md regions|pipe.run|multi-domain" >> $L
done
for c in "c2 4" "c4 1" "c1 1" "c3 4"; do
  echo "== WITCH_TIMING $c" >> $L
  HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py $c 2>&1 | grep -E "witch timing|pipe.run" >> $L
done
cat $L
