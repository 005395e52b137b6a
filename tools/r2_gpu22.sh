#!/bin/bash
# Round-2 GPU call 22/23 (N GPUs): strong scaling of config c5 (100,000 fragmentary queries x 315 HMMs): a step = a quarter of the
# query set, split over the ranks by sharding.partition_queries.
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
L=gpurun_out/r2_gpu22_$N.log
: > $L
t0=$(date +%s)
if [ "$N" = "1" ]; then R="python bench.py"; else R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py"; fi
timeout 1300 $R --gpus $N --config c5 --strong --slabs 4 --steps 4 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_c5_strong_${N}gpu.json 2> gpurun_out/r02_bench_c5_strong_${N}gpu.err; echo "c5 strong N=$N rc=$? wall $(( $(date +%s) - t0 )) s" >> $L
python -c "
import json
try:
    d=json.load(open('gpurun_out/r02_bench_c5_strong_${N}gpu.json')); print('N=$N', d['scaling'], round(d['value'],1), 'GCUPS', round(d['queries_per_s'],1), 'q/s e2e', round(d['e2e']['value'],1), round(d['ms_per_step'],1), d['clocks'], d['workload_stats']['queries_processed'])
except Exception as ex: print('FAILED', ex)
" >> $L
tail -3 gpurun_out/r02_bench_c5_strong_${N}gpu.err >> $L
cat $L
