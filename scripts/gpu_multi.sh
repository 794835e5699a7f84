#!/bin/bash
# 2-GPU sanity: weak scaling of the default workload and the row-band sharded 4K frame
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu rc=$?"; tail -1 gpurun_out/bench_2gpu.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload cfg4 > gpurun_out/bench_cfg4_2gpu.json 2> gpurun_out/bench_cfg4_2gpu.err; echo "cfg4 2gpu rc=$?"; tail -1 gpurun_out/bench_cfg4_2gpu.json | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_2gpu.json 2>&1; echo "ref 2gpu rc=$?"; tail -1 gpurun_out/bench_ref_2gpu.json | cut -c1-200
