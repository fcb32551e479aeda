#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 600 python bench.py --workload 'configs[2]' --no-cpu-baseline --no-pixels 2>gpurun_out/t_c2.err | tail -1 > gpurun_out/t_c2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/t_c2.json'))
print('configs[2]', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
PY
