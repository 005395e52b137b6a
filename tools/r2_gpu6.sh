#!/bin/bash
# Round-2 GPU call 6: two-level E scan -- parity tests and multi-domain timing per shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu6.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -3 gpurun_out/r2_pytest6.log >> $L
run() { # cfg slabs spread
  echo "== $1 spread=$3" >> $L
  WITCH_MD_SPREAD=$3 HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py $1 $2 2>&1 | grep -E "md regions|pipe.run|single-domain" >> $L
}
run c2 4 0; run c1 1 0; run c1 1 8; run c4 1 0; run c4 1 2
cat $L
