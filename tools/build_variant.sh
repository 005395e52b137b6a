#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]  -> tools/bin/libwitch_<name>.so (perf experiments; PERF_LIB selects it)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -warn-spills -shared -Xcompiler -fPIC -cudart shared \
  "$@" -o tools/bin/libwitch_$name.so witch_b200/csrc/witch_abi.cu witch_b200/csrc/hmm_profile.cpp 2>&1 | grep -vE "warning #|^\s*$|\^|Remark|was set but|declared but|detected during|instantiation of" || true
ls -la tools/bin/libwitch_$name.so
