#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/k_tests.log
timeout 300 python tools/timeline_frame.py > gpurun_out/k_timeline.txt 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/k_bench.json
cut -c1-400 gpurun_out/k_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/k_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','windows_ms_per_step','e2e','roofline')})
PY
