#!/bin/bash
# A/B of experimental library builds on cfg3 r=7 (each run under its own timeout), interleaved twice
mkdir -p gpurun_out
for rep in 1 2; do
for v in ${VARIANTS}; do
  echo -n "$v: "; ASW_B200_LIB=$PWD/stereo_matchin_b200/libasw_b200_$v.so timeout 120 python scripts/profile_run.py cfg3 7 0 2 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k: round(d[k],3) for k in ('raw_ms','supp_ms','vagg_mean_ms','hagg_mean_ms','wta_ms','total_ms')})"
done
done
for v in ${VARIANTS}; do
  echo "parity $v"; ASW_B200_LIB=$PWD/stereo_matchin_b200/libasw_b200_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=200 -k "edge or golden or band or cfg3" 2>&1 | tail -2
done
