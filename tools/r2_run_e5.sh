#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "bank_shift or clip or steady or graph or api or offload" --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-pixels 2>gpurun_out/d_bench.err | tail -1 > gpurun_out/e_bench.json
python - <<'PY'
import json,sys
d=json.load(open('gpurun_out/e_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','windows_ms_per_step','gpu_launches')}, d['e2e']['value'])
PY
python tools/timeline_frame.py > gpurun_out/r2_timeline_now.txt 2>&1; grep -v "^Exception\|Traceback\|  File \|Attribute" gpurun_out/r2_timeline_now.txt | tail -12
