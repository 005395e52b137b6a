set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/r1_pytest_gpu_final.log
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_bench_c2_v6.json 2> gpurun_out/r1_bench_c2_v6.err
T="python bench.py --max-queries 640 --max-hmms 48 --steps 1 --warmup 0 --no-cpu-baseline"
$T > gpurun_out/r1_trunc_v6.json 2> gpurun_out/r1_trunc_v6.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_launches_trunc.csv $T > gpurun_out/r01b_ncu_launches.log 2>&1
ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section MemoryWorkloadAnalysis --clock-control none -k regex:"wave_kernel|mh_parser" -c 10 -o gpurun_out/prof_r1b_sections $T > gpurun_out/r01b_ncu_sections.log 2>&1
ls -la gpurun_out/ | tail -8
tail -3 gpurun_out/r1_pytest_gpu_final.log; cut -c1-400 gpurun_out/r1_bench_c2_v6.json
