#!/bin/bash
# Round-2 GPU call 28: last sanity check of the committed default build (fine envelope classes, merged align classes).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu28.log
: > $L
timeout 600 python -m pytest tests -m gpu -q -x -k "scores_weights or edge or mirror or c1" > gpurun_out/r2_pytest28.log 2>&1; echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_pytest28.log >> $L
timeout 600 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/r02_bench_c2_1gpu_v7.json 2> gpurun_out/r02_bench_c2_1gpu_v7.err; echo "bench rc=$?" >> $L
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c2_1gpu_v7.json')); print('c2', round(d['value'],1), 'GCUPS e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), {k[:24]:(round(v['ms']/d['steps'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
" >> $L 2>&1
cat $L
