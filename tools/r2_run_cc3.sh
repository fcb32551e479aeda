#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cc_t' -c 12 --csv --log-file gpurun_out/cc_launches.csv python tools/bench_kernels.py --once > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/cc_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:12]: print(r[4][:70], r[-1])
PY
