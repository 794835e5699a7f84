#!/bin/bash
# 8-GPU evidence run (charged 8x): the bench line at N = 8 (cfg3 replicas + also.cfg4_strong through NCCL ranks + also.cfg5_batch)
# and the one-process asw_multi_* path on the cfg4 frame at 1 and N devices (equality with the 1-GPU map checked in the run).
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-300
timeout 300 python scripts/multi_probe.py 1 $N > gpurun_out/multi_probe_${N}.jsonl 2> gpurun_out/multi_probe_${N}.err; echo "multi_probe rc=$?"; cat gpurun_out/multi_probe_${N}.jsonl
