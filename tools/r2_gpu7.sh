#!/bin/bash
# Round-2 GPU call 7: two-lane batches of the multi-domain branch; live-reference counts; parity tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu7.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -3 gpurun_out/r2_pytest7.log >> $L
timeout 600 python -m pytest tests/test_live_reference.py -m gpu -q -s 2>&1 | grep -E "reported pairs|passed|failed" > gpurun_out/r02_live_reference.txt
cat gpurun_out/r02_live_reference.txt >> $L
run() { # cfg slabs spread
  echo "== $1 spread=$3" >> $L
  WITCH_MD_SPREAD=$3 HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py $1 $2 2>&1 | grep -E "md regions|pipe.run|single-domain" >> $L
}
run c4 1 0; run c4 1 1; run c1 1 0; run c1 1 8
cat $L
