#!/bin/bash
# timing-only A/B of experimental library builds on cfg3 r=7: VARIANTS="a b c" REPS=2 bash scripts/gpu_ab.sh
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do
for v in ${VARIANTS}; do
  echo -n "$v: "; ASW_B200_LIB=$PWD/stereo_matchin_b200/libasw_b200_$v.so timeout 120 python scripts/profile_run.py cfg3 7 0 2 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(d[k],3) for k in ('vagg_mean_ms','hagg_mean_ms','total_ms')})"
done
done
