#!/bin/bash
# Prepared wave-kernel experiments (WITCH_WAVE_EXP bit mask): build every variant HERE (no GPU needed),
# then run the parity + throughput probe of each on a B200 in ONE gpurun call.
#   tools/r2_variants.sh build            -> tools/bin/libwitch_exp{0..7}.so   (git-ignored, travels with the gpurun snapshot)
#   gpurun --timeout 900 -- 'bash tools/r2_variants.sh run'   -> gpurun_out/r2_variants.log
# exp0 is the default kernel (its score dump is the reference the others are compared with: reported sets must be
# identical and max|dscore| ~1e-5 bits, tools/gpu_perf_c2.py prints both).
set -e
cd "$(dirname "$0")/.."
case "$1" in
build)
  for v in 0 1 2 3 4 5 6 7; do bash tools/build_variant.sh exp$v -DWITCH_WAVE_EXP=$v > /dev/null 2>&1 & done; wait
  # the two-items-per-warp envelope kernel (wave_pair_kernel.cuh): 3 / 4 CTAs per SM, with and without the elect.sync issue
  bash tools/build_variant.sh pair -DWITCH_WAVE_PAIR=1 > /dev/null 2>&1 &
  bash tools/build_variant.sh pair_e1 -DWITCH_WAVE_PAIR=1 -DWITCH_WAVE_EXP=1 > /dev/null 2>&1 &
  bash tools/build_variant.sh pair_mb4 -DWITCH_WAVE_PAIR=1 -DWITCH_PAIR_MINB=4 > /dev/null 2>&1 &
  bash tools/build_variant.sh pair_e7 -DWITCH_WAVE_PAIR=1 -DWITCH_WAVE_EXP=7 > /dev/null 2>&1 & wait
  ls -la tools/bin/libwitch_exp?.so tools/bin/libwitch_pair*.so ;;
run)
  mkdir -p gpurun_out
  : > gpurun_out/r2_variants.log
  PERF_LIB=tools/bin/libwitch_exp0.so python tools/gpu_perf_c2.py 640 48 base >> gpurun_out/r2_variants.log 2>&1
  for v in exp1 exp2 exp4 exp7 pair pair_e1 pair_mb4 pair_e7; do
    PERF_LIB=tools/bin/libwitch_$v.so timeout 300 python tools/gpu_perf_c2.py 640 48 $v >> gpurun_out/r2_variants.log 2>&1 || echo "$v FAILED rc=$?" >> gpurun_out/r2_variants.log
  done
  # full GPU parity tests against the pair kernel (the default library is covered by `pytest -m gpu` itself)
  timeout 900 python tools/run_tests_with_lib.py tools/bin/libwitch_pair.so > gpurun_out/r2_pair_pytest.log 2>&1 || echo "pair parity tests FAILED" >> gpurun_out/r2_variants.log
  tail -n 3 gpurun_out/r2_pair_pytest.log >> gpurun_out/r2_variants.log
  cat gpurun_out/r2_variants.log ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac
