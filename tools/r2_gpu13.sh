#!/bin/bash
# Round-2 GPU call 13: pair parser for the C=4 classes -- parity tests; chain1 variant of generation 6; c4 / c1 / c3 bench lines and c4 host phases.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu13.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_pytest13.log >> $L
timeout 300 python tools/gpu_perf_c2.py 640 48 base 2>&1 | grep -E "^\[|vs base|rror" >> $L
PERF_LIB=tools/bin/libwitch_chain1.so timeout 300 python tools/gpu_perf_c2.py 640 48 chain1 2>&1 | grep -E "^\[|vs base|rror" >> $L
B="python bench.py --no-cpu-baseline"
timeout 600 $B --config c4 --slabs 1 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_1gpu.json 2> gpurun_out/r02_bench_c4_1gpu.err
timeout 600 $B --config c1 --slabs 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1_1gpu.json 2> gpurun_out/r02_bench_c1_1gpu.err
timeout 900 $B --config c3 --slabs 4 --steps 4 --warmup 2 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err
for f in c4 c1 c3; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r02_bench_${f}_1gpu.json')); print('$f', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), {k[:24]:(round(v['ms']),round(v['gcells_per_s'])) for k,v in d['roofline']['kernels'].items()})
except Exception as ex: print('$f FAILED', ex)
" >> $L; done
echo "== WITCH_TIMING c4" >> $L
HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py c4 1 2>&1 | grep -E "witch timing|pipe.run" | tail -8 >> $L
rm -f gpurun_out/scores_*.npz
cat $L
