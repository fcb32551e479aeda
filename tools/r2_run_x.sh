#!/bin/bash
# state check after the container was re-created: full GPU suite, smoke, default bench, kernel stress shapes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider ) > gpurun_out/x_tests.log 2>&1; tail -6 gpurun_out/x_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/x_smoke.log 2>&1; tail -2 gpurun_out/x_smoke.log
timeout 600 python bench.py 2>gpurun_out/x_bench.err | tail -1 > gpurun_out/x_bench.json; tail -3 gpurun_out/x_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/x_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','windows_ms_per_step')}, d['e2e'], d['roofline'], d.get('roofline_hbm'))
PY
timeout 300 python tools/bench_kernels.py > gpurun_out/x_kernels.log 2>&1; tail -20 gpurun_out/x_kernels.log
