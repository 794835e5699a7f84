#!/bin/bash
for nw in 8 4; do echo -n "ASW_V_NW=$nw "; ASW_V_NW=$nw timeout 120 python scripts/profile_run.py cfg3 7 0 2 | python -c "import json,sys; d=json.load(sys.stdin); print(' V %.3f H %.3f total %.2f'%(d['vagg_mean_ms'],d['hagg_mean_ms'],d['total_ms']))"; done
ASW_V_NW=4 timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -3
ASW_V_NW=8 timeout 600 python -m pytest tests -m gpu -q -x --timeout=300 2>&1 | tail -3
