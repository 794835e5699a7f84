#!/bin/bash
# N-GPU sanity (N = $1, default 2): weak scaling of the default workload and the sharded 4K frame (2-D grid and row bands only)
N=${1:-2}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -1 gpurun_out/$name.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus')}, d['config']['sharding'], 'e2e', round(d['e2e']['value']))
except Exception as e: print('no json line', e)
"; tail -3 gpurun_out/$name.err; }
run bench_${N}gpu --steps 3 --warmup 3 --no-cpu-baseline
run bench_cfg4_${N}gpu --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline
run bench_cfg4_bands_${N}gpu --steps 3 --warmup 3 --workload cfg4 --bands-only --no-cpu-baseline
