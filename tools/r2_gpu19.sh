#!/bin/bash
# Round-2 GPU call 19: parity tests incl. the parser-class test, the default bench line (with the CPU reference leg), c1 / c3 lines with the final parser.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu19.log
: > $L
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?" >> $L
tail -3 gpurun_out/r2_pytest19.log >> $L
timeout 900 python bench.py > gpurun_out/r02_bench_c2_1gpu_v3.json 2> gpurun_out/r02_bench_c2_1gpu_v3.err; echo "bench rc=$?" >> $L
B="python bench.py --no-cpu-baseline"
timeout 600 $B --config c1 --slabs 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1_1gpu.json 2> gpurun_out/r02_bench_c1_1gpu.err
timeout 900 $B --config c3 --slabs 4 --steps 4 --warmup 2 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err
for f in c2_1gpu_v3 c1_1gpu c3_1gpu; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), {k[:24]:(round(v['ms']),round(v['gcells_per_s'])) for k,v in d['roofline']['kernels'].items()}, d['e2e'].get('host_ms_per_step'), d['roofline']['bound'], round(d['roofline']['frac'],3), d['cpu_baseline'] and round(d['cpu_baseline']['value'],3))
except Exception as ex: print('$f FAILED', ex)
" >> $L; done
cat $L
