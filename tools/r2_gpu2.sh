#!/bin/bash
# Round-2 GPU call 2: parity tests with the multi-domain branch on the device, smoke, host-phase timing, bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke2.log
WITCH_TIMING=1 timeout 600 python tools/gpu_hosttime.py c2 4 > gpurun_out/r2_hosttime2.log 2>&1
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_bench2_n1.json 2> gpurun_out/r2_bench2_n1.err; echo "bench rc=$?" >> gpurun_out/r2_bench2_n1.err
tail -15 gpurun_out/r2_pytest2.log; cat gpurun_out/r2_smoke2.log; tail -22 gpurun_out/r2_hosttime2.log; head -c 600 gpurun_out/r2_bench2_n1.json; tail -3 gpurun_out/r2_bench2_n1.err
