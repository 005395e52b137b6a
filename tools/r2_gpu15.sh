#!/bin/bash
# Round-2 GPU call 15: C=13 hybrid parser (default) vs the C=9 / 192-thread experiment; ncu --set full of the new parser; full c2 bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu15.log
: > $L
timeout 300 python tools/gpu_perf_c2.py 640 48 base 2>&1 | grep -E "^\[|vs base|rror" >> $L
WITCH_PARSER_C9=1 timeout 300 python tools/gpu_perf_c2.py 640 48 c9 2>&1 | grep -E "^\[|vs base|rror" >> $L
T="python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 1 --warmup 0 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_parser2 --launch-count 1 -f -o gpurun_out/prof_r02_parser_c13 $T > gpurun_out/r02_ncu_parser_c13.log 2>&1
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c2_1gpu_v2.json 2> gpurun_out/r02_bench_c2_1gpu_v2.err; echo "bench rc=$?" >> $L
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c2_1gpu_v2.json')); print('c2', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), d['ms_per_step'], {k[:24]:(round(v['ms']),round(v['gcells_per_s'])) for k,v in d['roofline']['kernels'].items()})
" >> $L 2>&1
rm -f gpurun_out/scores_*.npz
cat $L
