#!/bin/bash
# Round-2 GPU call 26: align stage with merged length classes -- parity tests that align, truncated bench (align share), full bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu26.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x -k "scores_weights or c1 or properties or mirror or device_pipeline or next_rows or limits" > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_pytest26.log >> $L
python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_c2trunc_v2.json 2> /dev/null
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c2_1gpu_v5.json 2> gpurun_out/r02_bench_c2_1gpu_v5.err; echo "bench rc=$?" >> $L
for f in c2trunc_v2 c2_1gpu_v5; do python -c "
import json
d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', round(d['value'],1), 'GCUPS e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), {k[:24]:(round(v['ms']/d['steps'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
" >> $L 2>&1; done
cat $L
