#!/bin/bash
# multi-GPU run on one box: N=${N:-2}.  Headline (cfg3 replicas, weak) with the also.cfg4_strong / also.cfg5_batch blocks,
# and cfg4 as the headline (strong scaling, halo exchange).
mkdir -p gpurun_out
N=${N:-2}
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -1 gpurun_out/$name.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d[k] for k in ('value', 'ms_per_step', 'n_gpus', 'scaling')}, 'e2e', round(d['e2e']['value'], 1))
for k, v in (d.get('also') or {}).items():
    print(' ', k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a in ('ms_per_frame', 'Mpix_disp_per_s', 'equals_1gpu', 'pairs_per_s', 'seconds_for_1024_pairs_extrapolated')})
"; }
nvidia-smi -L | head -8
run bench_${N}gpu --steps 5 --warmup 3
run bench_cfg4_${N}gpu --steps 3 --warmup 3 --workload cfg4 --no-cpu-baseline
