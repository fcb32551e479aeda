#!/bin/bash
mkdir -p gpurun_out
for w in 2 3 4; do
  timeout 900 python bench.py --workload "configs[$w]" --steps 20 --warmup 3 --no-cpu-baseline --no-pixels --sweep-clips 48 2>gpurun_out/u_bench_$w.err | tail -1 > gpurun_out/u_bench_cfg$w.json
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/u_bench_cfg$w.json'))
    print($w, {k:d.get(k) for k in ('value','ms_per_step','frame_objects_per_s','unit')}, (d.get('e2e') or {}).get('value'))
except Exception as e:
    print($w, 'failed', e); print(open('gpurun_out/u_bench_$w.err').read()[-1500:])
PY
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/u_ref.err | tail -1 > gpurun_out/u_bench_reference.json
cut -c1-400 gpurun_out/u_bench_reference.json
