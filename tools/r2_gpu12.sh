#!/bin/bash
# Round-2 GPU call 12: generation-6 parser as the default -- parity tests, the default bench line (full c2), parser generations on c4.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu12.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?" >> $L
tail -3 gpurun_out/r2_pytest12.log >> $L
timeout 900 python bench.py > gpurun_out/r02_bench_c2_1gpu.json 2> gpurun_out/r02_bench_c2_1gpu.err; echo "bench rc=$?" >> $L
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c2_1gpu.json')); print('c2', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), d['ms_per_step'], {k[:24]:(round(v['ms']),round(v['gcells_per_s'])) for k,v in d['roofline']['kernels'].items()}, d['cpu_baseline'])
" >> $L 2>&1
WITCH_PARSER=1 timeout 400 python tools/gpu_perf_c2.py 1500 64 base c4 2>&1 | grep -E "^\[|vs base|rror" >> $L
for g in 2 3; do WITCH_PARSER=$g timeout 300 python tools/gpu_perf_c2.py 1500 64 gen$g c4 2>&1 | grep -E "^\[|vs base|rror" >> $L; done
rm -f gpurun_out/scores_*.npz
cat $L
