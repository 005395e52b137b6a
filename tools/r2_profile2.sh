#!/bin/bash
# Round-2 profile capture (final kernels): bench line + ncu launch list of the truncated bench command, --set full captures
# of the parser / envelope / align launches of that command and of the multi-domain kernels on a c4-like sample.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 2 --warmup 1 --no-cpu-baseline"
$T > gpurun_out/r02_bench_c2trunc_same_command.json 2> gpurun_out/r02_bench_c2trunc.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_ncu_launch_list_c2trunc.csv $T > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mh_parser --launch-count 5 -f -o gpurun_out/prof_r02_parser $T > gpurun_out/r02_ncu_parser.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wave_kernel --launch-count 10 -f -o gpurun_out/prof_r02_wave $T > gpurun_out/r02_ncu_wave.log 2>&1
M="python tools/gpu_perf_c2.py 1500 64 c4md c4"
$M > gpurun_out/r02_c4md_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:md_ --launch-count 3 -f -o gpurun_out/prof_r02_md_c4 $M > gpurun_out/r02_ncu_md_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r02_c4md_plain.log; head -c 400 gpurun_out/r02_bench_c2trunc_same_command.json
