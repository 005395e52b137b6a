#!/bin/bash
# Round-2 final check on a B200: build entry smoke, GPU parity tests, the default bench line, the reference arm (short).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_final.log
: > $L
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?" >> $L
tail -1 gpurun_out/r2_final_smoke.log >> $L
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?" >> $L
tail -2 gpurun_out/r2_final_pytest.log >> $L
timeout 900 python bench.py > gpurun_out/r02_bench_c2_1gpu_final.json 2> gpurun_out/r02_bench_c2_1gpu_final.err; echo "bench rc=$?" >> $L
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c2_reference_arm.json 2> gpurun_out/r02_bench_c2_reference_arm.err; echo "reference arm rc=$?" >> $L
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c2_1gpu_final.json')); print('c2', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['roofline']['bound'], round(d['roofline']['frac'],3), d['cpu_baseline']['value'])
r=json.load(open('gpurun_out/r02_bench_c2_reference_arm.json')); print('reference arm', r['value'], r['unit'], r['cpu_baseline']['cores'], 'cores')
" >> $L 2>&1
cat $L
