#!/bin/bash
# Round-2 GPU call 8 (after the container was re-created): parity tests, the default bench line, the profile set of the
# truncated bench command (launch list + --set full captures), host phases, and the 16-bit stored-row variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu8.log
: > $L
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log
tail -3 gpurun_out/r2_pytest8.log >> $L
timeout 900 python bench.py > gpurun_out/r02_bench_c2_1gpu.json 2> gpurun_out/r02_bench_c2_1gpu.err; echo "bench rc=$?" >> $L
T="python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 2 --warmup 1 --no-cpu-baseline"
$T > gpurun_out/r02_bench_c2trunc_same_command.json 2> gpurun_out/r02_bench_c2trunc.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_ncu_launch_list_c2trunc.csv $T > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_parser --launch-count 4 -f -o gpurun_out/prof_r02_parser $T > gpurun_out/r02_ncu_parser.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_kernel --launch-count 8 -f -o gpurun_out/prof_r02_wave $T > gpurun_out/r02_ncu_wave.log 2>&1
echo "== WITCH_TIMING c2" >> $L
HOSTTIME_REPS=2 WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py c2 4 > gpurun_out/r02_host_phases.txt 2>&1
grep -E "pipe.run" gpurun_out/r02_host_phases.txt >> $L
: > gpurun_out/r2_variants8.log
timeout 300 python tools/gpu_perf_c2.py 640 48 base >> gpurun_out/r2_variants8.log 2>&1
for v in row16_1 row16_2; do
  PERF_LIB=tools/bin/libwitch_$v.so timeout 300 python tools/gpu_perf_c2.py 640 48 $v >> gpurun_out/r2_variants8.log 2>&1 || echo "$v FAILED rc=$?" >> gpurun_out/r2_variants8.log
done
grep -vE "Warning|warn" gpurun_out/r2_variants8.log >> $L
ls -la gpurun_out/*.ncu-rep >> $L
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c2_1gpu.json')); print('c2', round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), d['ms_per_step'], {k[:12]:(round(v['ms']),round(v['gcells_per_s'])) for k,v in d['roofline']['kernels'].items()}, d['cpu_baseline'])
" >> $L 2>&1
cat $L
