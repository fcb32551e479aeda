#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py 2>gpurun_out/b_bench.err | tail -1 > gpurun_out/b_bench.json; tail -2 gpurun_out/b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','windows_ms_per_step')}, d['e2e']['value'], d['roofline']['frac'])
PY
python tools/timeline_frame.py > gpurun_out/r2_timeline_now.txt 2>&1; grep -v "^Exception\|Traceback\|File \|Attribute" gpurun_out/r2_timeline_now.txt | sed -n 4,16p; grep "frame span" gpurun_out/r2_timeline_now.txt
