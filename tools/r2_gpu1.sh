#!/bin/bash
# Round-2 GPU call 1: parity tests, A/B of the prepared envelope-kernel variants on truncated c2, first bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
: > gpurun_out/r2_variants.log
PERF_LIB=tools/bin/libwitch_exp0.so timeout 300 python tools/gpu_perf_c2.py 640 48 base >> gpurun_out/r2_variants.log 2>&1
for v in exp7 pair pair_e7 row16a row16b row16a_e7; do
  PERF_LIB=tools/bin/libwitch_$v.so timeout 300 python tools/gpu_perf_c2.py 640 48 $v >> gpurun_out/r2_variants.log 2>&1 || echo "$v FAILED rc=$?" >> gpurun_out/r2_variants.log
done
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2_bench_n1.err
tail -3 gpurun_out/r2_pytest.log; cat gpurun_out/r2_variants.log; tail -c 1500 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
