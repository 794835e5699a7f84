#!/bin/bash
# One GPU session: smoke, parity tests, probes, a short bench.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
python -m pytest tests -m gpu -q --maxfail=15 -x --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/summary.txt
tail -30 gpurun_out/pytest_gpu.log
python scripts/ubench.py > gpurun_out/ubench.json 2>&1; echo "ubench rc=$?" | tee -a gpurun_out/summary.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_cfg3.json
tail -5 gpurun_out/bench_cfg3.err
