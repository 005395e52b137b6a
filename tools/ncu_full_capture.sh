#!/bin/bash
# ncu --set full captures of the longest parser and envelope launch of the truncated bench command (run under gpurun)
T="python bench.py --max-queries 640 --max-hmms 48 --steps 1 --warmup 0 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:mh_parser --launch-count 1 -f -o gpurun_out/prof_r1c_parser $T > gpurun_out/r01c_ncu_parser.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wave_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r1c_env $T > gpurun_out/r01c_ncu_env.log 2>&1
ls -la gpurun_out/prof_r1c_*
