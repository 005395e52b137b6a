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
  ls -la tools/bin/libwitch_exp?.so ;;
run)
  mkdir -p gpurun_out
  : > gpurun_out/r2_variants.log
  PERF_LIB=tools/bin/libwitch_exp0.so python tools/gpu_perf_c2.py 640 48 base >> gpurun_out/r2_variants.log 2>&1
  for v in 1 2 4 7; do
    PERF_LIB=tools/bin/libwitch_exp$v.so timeout 300 python tools/gpu_perf_c2.py 640 48 exp$v >> gpurun_out/r2_variants.log 2>&1 || echo "exp$v FAILED rc=$?" >> gpurun_out/r2_variants.log
  done
  cat gpurun_out/r2_variants.log ;;
*) echo "usage: $0 build|run"; exit 2 ;;
esac
