#!/bin/bash
# Round-2 GPU call 20 (N GPUs): the driver's launch line at N = $1 (weak scaling, default steps).
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r2_gpu20.log
: > $L
t0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 4 --warmup 4 > gpurun_out/r02_bench_c2_${N}gpu.json 2> gpurun_out/r02_bench_c2_${N}gpu.err; echo "weak N=$N rc=$? wall $(( $(date +%s) - t0 )) s" >> $L
python -c "
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c2_${N}gpu.json')); print('N=$N', d['scaling'], round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['clocks'])
except Exception as ex: print('FAILED', ex)
" >> $L
wc -l gpurun_out/r02_bench_c2_${N}gpu.json >> $L
tail -3 gpurun_out/r02_bench_c2_${N}gpu.err >> $L
cat $L
