#!/bin/bash
# Round-end GPU session: parity tests, bench line, ncu launch list of the bench command, full ncu
# probes (the ncu --set full capture is a separate call: scripts/gpu_profile.sh).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/ubench.py > gpurun_out/ubench.json 2>&1; echo "ubench rc=$?"
[ -x scripts/taploop.bin ] && { timeout 200 scripts/taploop.bin 1 > gpurun_out/taploop.jsonl 2>&1; timeout 60 scripts/taploop.bin 3 >> gpurun_out/taploop.jsonl 2>&1; timeout 60 scripts/taploop.bin 4 >> gpurun_out/taploop.jsonl 2>&1; }; echo "taploop rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench rc=$?"; cat gpurun_out/bench_cfg3.json
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "bench launch list rc=$?"
for wl in cfg2 cfg5 cfg4; do timeout 300 python scripts/profile_run.py $wl 7 0 3 > gpurun_out/stage_$wl.json 2>&1; tail -1 gpurun_out/stage_$wl.json; done
timeout 600 python bench.py --impl reference --steps 2 > gpurun_out/bench_reference.json 2>&1; tail -1 gpurun_out/bench_reference.json
