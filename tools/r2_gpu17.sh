#!/bin/bash
# Round-2 GPU call 17 (2 GPUs): the driver's launch line at N=2 (weak scaling, default steps) and a strong-scaling line of the same slabs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=gpurun_out/r2_gpu17.log
: > $L
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2"
timeout 900 $R --steps 4 --warmup 4 > gpurun_out/r02_bench_c2_2gpu.json 2> gpurun_out/r02_bench_c2_2gpu.err; echo "weak rc=$?" >> $L
timeout 900 $R --steps 4 --warmup 4 --strong > gpurun_out/r02_bench_c2_2gpu_strong.json 2> gpurun_out/r02_bench_c2_2gpu_strong.err; echo "strong rc=$?" >> $L
for f in c2_2gpu c2_2gpu_strong; do python -c "
import json
try:
    d=json.load(open('gpurun_out/r02_bench_$f.json')); print('$f', d['scaling'], round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['clocks'])
except Exception as ex: print('$f FAILED', ex)
" >> $L; done
tail -3 gpurun_out/r02_bench_c2_2gpu.err >> $L
cat $L
