#!/bin/bash
for hs in 0 1; do echo "ASW_H_SPLIT=$hs"; ASW_H_SPLIT=$hs timeout 120 python scripts/profile_run.py cfg3 7 0 2 | python -c "import json,sys; d=json.load(sys.stdin); print(' V %.3f H %.3f total %.2f'%(d['vagg_mean_ms'],d['hagg_mean_ms'],d['total_ms']))"; done
timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -3
