#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/m_tests.log
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/m_bench.err | tail -1 > gpurun_out/m_bench.json
tail -5 gpurun_out/m_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/m_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','windows_ms_per_step','e2e','e2e_from_pixels','whole_clip')})
PY
