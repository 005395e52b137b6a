#!/bin/bash
# Round-2 GPU call 11: ncu --set full of the generation-6 parser (one launch) on the truncated bench command.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python bench.py --max-queries 640 --max-hmms 48 --slabs 1 --steps 1 --warmup 0 --no-cpu-baseline"
WITCH_PARSER=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_parser2 --launch-count 1 -f -o gpurun_out/prof_r02_parser_gen6 $T > gpurun_out/r02_ncu_parser_gen6.log 2>&1
ls -la gpurun_out/
