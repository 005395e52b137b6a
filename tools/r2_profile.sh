#!/bin/bash
# Round-2 measurement call: bench lines of every named shape, ncu launch list and --set full captures of the truncated
# bench command (run under gpurun; results land in gpurun_out/, summaries are made here with tools/ncu_summary.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline"
python bench.py --steps 8 --warmup 4 > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err
$B --config c1 --slabs 1 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1_n1.json 2> gpurun_out/r02_bench_c1_n1.err
$B --config c4 --slabs 1 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_n1.json 2> gpurun_out/r02_bench_c4_n1.err
$B --config c3 --slabs 4 --steps 4 --warmup 3 > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/r02_bench_c3_n1.err
$B --config c5 --slabs 16 --steps 8 --warmup 3 > gpurun_out/r02_bench_c5_100k_n1.json 2> gpurun_out/r02_bench_c5_100k_n1.err
T="python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 2 --warmup 1 --no-cpu-baseline"
$T > gpurun_out/r02_bench_c2trunc_same_command.json 2> gpurun_out/r02_bench_c2trunc.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launch_list_c2trunc.csv $T > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mh_parser --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r02_parser $T > gpurun_out/r02_ncu_parser.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wave_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_r02_env $T > gpurun_out/r02_ncu_env.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:md_region --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r02_md $T > gpurun_out/r02_ncu_md.log 2>&1
ls -la gpurun_out/ | tail -20
for f in c2_n1 c1_n1 c4_n1 c3_n1 c5_100k_n1; do python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as ex: print('$f FAILED', ex)
"; done
