#!/bin/bash
# Builds tools/sim/libwitch_sim.so: the library's own sources compiled with g++ against the SIMT simulator (TEST TOOLING).
# usage: tools/sim/build_sim.sh [name] [extra -D flags]   -> tools/sim/libwitch_<name>.so (default name: sim)
set -e
cd "$(dirname "$0")/../.."
name=${1:-sim}; shift || true
g++ -std=c++17 -O1 -g -fPIC -shared -DWITCH_HOST_SIM -Wno-unknown-pragmas -Itools/sim/shim "$@" \
    -x c++ witch_b200/csrc/witch_abi.cu -x c++ witch_b200/csrc/hmm_profile.cpp tools/sim/simt.cpp -o tools/sim/libwitch_$name.so
ls -la tools/sim/libwitch_$name.so
