#!/bin/bash
# A/B of one library build under different environment knobs: LIBV=name KNOBS="ASW_H_LDG=0 ASW_H_LDG=1" bash scripts/gpu_ab2.sh
mkdir -p gpurun_out
export ASW_B200_LIB=$PWD/stereo_matchin_b200/libasw_b200_${LIBV}.so
for rep in $(seq 1 ${REPS:-2}); do
for k in ${KNOBS}; do
  echo -n "$k: "; env $k timeout 120 python scripts/profile_run.py cfg3 7 0 2 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(d[k],3) for k in ('vagg_mean_ms','hagg_mean_ms','total_ms')})"
done
done
