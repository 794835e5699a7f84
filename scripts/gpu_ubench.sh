#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ubench.py > gpurun_out/ubench.json 2>gpurun_out/ubench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/ubench.json'))
for k,v in d['lds_sm_cycles_per_warp_instruction'].items(): print(k, v)
"
