#!/bin/bash
for dbg in 0 4 8 12; do echo "ASW_DBG=$dbg"; ASW_DBG=$dbg timeout 120 python scripts/profile_run.py cfg3 3 0 2 | python -c "import json,sys; d=json.load(sys.stdin); print(' V %.3f H %.3f total %.2f'%(d['vagg_mean_ms'],d['hagg_mean_ms'],d['total_ms']))"; done
