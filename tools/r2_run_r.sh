#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-pixels 2>gpurun_out/r_bench.err | tail -1 > gpurun_out/r_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','windows_ms_per_step')}, d['e2e']['value'], d['e2e']['windows_ms_per_step'])
PY
tail -3 gpurun_out/r_bench.err
