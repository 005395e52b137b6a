#!/bin/bash
# Round-2 GPU call 4: parity tests (default parser and WITCH_PARSER=4), multi-domain timing per shape, bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu4.log
: > $L
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -5 gpurun_out/r2_pytest4.log >> $L
WITCH_PARSER=4 timeout 1200 python -m pytest tests -m gpu -q -k "scores_weights or properties or device_pipeline or c1_cuda or live" > gpurun_out/r2_pytest4_gen4.log 2>&1; echo "pytest gen4 rc=$?" >> gpurun_out/r2_pytest4_gen4.log
tail -5 gpurun_out/r2_pytest4_gen4.log >> $L
for c in "c2 4" "c4 1" "c1 1"; do
  echo "== WITCH_TIMING $c" >> $L
  HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py $c 2>&1 | grep -E "witch timing|pipe.run" >> $L
done
B="python bench.py --no-cpu-baseline"
$B --steps 4 --warmup 4 > gpurun_out/r2b_bench_c2_n1.json 2> gpurun_out/r2b_bench_c2_n1.err
WITCH_PARSER=4 $B --steps 4 --warmup 4 > gpurun_out/r2b_bench_c2_n1_gen4.json 2> gpurun_out/r2b_bench_c2_n1_gen4.err
$B --config c4 --slabs 1 --steps 3 --warmup 3 > gpurun_out/r2b_bench_c4_n1.json 2> gpurun_out/r2b_bench_c4_n1.err
$B --config c1 --slabs 1 --steps 5 --warmup 3 > gpurun_out/r2b_bench_c1_n1.json 2> gpurun_out/r2b_bench_c1_n1.err
for f in c2_n1 c2_n1_gen4 c4_n1 c1_n1; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r2b_bench_$f.json')); print('$f', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), d['ms_per_step'], {k[:12]:(round(v['ms']),round(v['gcells_per_s'])) for k,v in d['roofline']['kernels'].items()})
except Exception as ex: print('$f FAILED', ex)
" >> $L; done
cat $L
