#!/bin/bash
N=8
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -1 gpurun_out/$name.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['config']['sharding'], 'e2e ms', round(d['e2e']['ms_per_step'], 2))
except Exception as e: print('no json line', e)
"; }
run cfg4_8gpu_grid --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline
run cfg4_8gpu_bands --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline --bands-only
run cfg3_8gpu --steps 3 --warmup 3 --no-cpu-baseline
