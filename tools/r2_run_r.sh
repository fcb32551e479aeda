#!/bin/bash
mkdir -p gpurun_out
for t in 120 296; do
VLS_TUNING="gemm_bn64_below=$t" timeout 300 python tools/timeline_frame.py > gpurun_out/r_timeline_$t.txt 2>&1
grep -E "frame span" gpurun_out/r_timeline_$t.txt
VLS_TUNING="gemm_bn64_below=$t" timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-pixels 2>gpurun_out/r_bench.err | tail -1 > gpurun_out/r_bench_$t.json
python - <<PY
import json
d=json.load(open('gpurun_out/r_bench_$t.json'))
print($t, {k:d.get(k) for k in ('value','ms_per_step','windows_ms_per_step')}, d['e2e']['value'], d['e2e']['windows_ms_per_step'])
PY
done
