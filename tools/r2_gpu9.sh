#!/bin/bash
# Round-2 profile set (GPU calls 9 and 18): the profile set of the truncated bench command with outputs that fit gpurun's 64 MiB return limit
# (launch list, --set full captures of 2 parser and 4 wave launches, host phases).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu9.log
: > $L
T="python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 2 --warmup 1 --no-cpu-baseline"
$T > gpurun_out/r02_bench_c2trunc_same_command.json 2> gpurun_out/r02_bench_c2trunc.err; echo "trunc bench rc=$?" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_ncu_launch_list_c2trunc.csv $T > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_parser --launch-count 2 -f -o gpurun_out/prof_r02_parser $T > gpurun_out/r02_ncu_parser.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_kernel --launch-count 4 -f -o gpurun_out/prof_r02_wave $T > gpurun_out/r02_ncu_wave.log 2>&1
HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py c2 4 > gpurun_out/r02_host_phases.txt 2>&1
grep -E "pipe.run" gpurun_out/r02_host_phases.txt >> $L

ls -la gpurun_out/*.ncu-rep >> $L
du -sm gpurun_out >> $L
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 60 ]; then rm -f gpurun_out/prof_r02_wave.ncu-rep; echo "wave rep dropped (size)" >> $L; fi
cat $L
