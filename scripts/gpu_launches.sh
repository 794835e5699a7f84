#!/bin/bash
# ncu launch list (durations only) of one cfg3 frame (r=2) after the same command exited 0 without ncu
mkdir -p gpurun_out
timeout 120 python scripts/profile_run.py cfg3 2 0 1 > gpurun_out/profile_plain.json 2>gpurun_out/profile_plain.err && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_run.py cfg3 2 0 1 > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h={n:i for i,n in enumerate(rows[hi])}
for r in rows[hi+2:]:
    if len(r)>h['Metric Value']: print(r[h['Kernel Name']][:50], r[h['Grid Size']] if 'Grid Size' in h else '', r[h['Metric Value']])
PY
