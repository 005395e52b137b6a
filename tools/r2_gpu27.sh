#!/bin/bash
# Round-2 GPU call 27: work-list length classes merged up to 2,048 residues (envelope pass too) -- truncated and full bench lines, live-reference + golden tests.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu27.log
: > $L
python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_c2trunc_v3.json 2> /dev/null
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c2_1gpu_v6.json 2> gpurun_out/r02_bench_c2_1gpu_v6.err; echo "bench rc=$?" >> $L
for f in c2trunc_v3 c2_1gpu_v6; do python -c "
import json
d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', round(d['value'],1), 'GCUPS e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), {k[:24]:(round(v['ms']/d['steps'],1),v['launches']) for k,v in d['roofline']['kernels'].items()})
" >> $L 2>&1; done
timeout 600 python -m pytest tests -m gpu -q -x -k "scores_weights or live or c1 or properties or edge" > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_pytest27.log >> $L
cat $L
