#!/bin/bash
# Round-2 GPU call 25: pooled query buffers -- parity tests that create many query sets, bench line (host phases of the e2e loop).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu25.log
: > $L
timeout 900 python -m pytest tests -m gpu -q -x -k "edge or parser_classes or mirror or device_pipeline or properties" > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_pytest25.log >> $L
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_c2_1gpu_v4.json 2> gpurun_out/r02_bench_c2_1gpu_v4.err; echo "bench rc=$?" >> $L
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c2_1gpu_v4.json')); print('c2', round(d['value'],1), 'GCUPS e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['e2e']['host_ms_per_step'])
" >> $L 2>&1
cat $L
