#!/bin/bash
mkdir -p gpurun_out
for t in 64 32; do
VLS_TUNING="attn_bal_min_tiles=$t" timeout 300 python tools/timeline_frame.py > gpurun_out/r_timeline_$t.txt 2>&1
grep -E "frame span" gpurun_out/r_timeline_$t.txt
grep -E "attn_fwd|attn_combine" gpurun_out/r_timeline_$t.txt | head -4
VLS_TUNING="attn_bal_min_tiles=$t" timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-pixels 2>gpurun_out/r_bench.err | tail -1 > gpurun_out/r_bench_$t.json
python - <<PY
import json
d=json.load(open('gpurun_out/r_bench_$t.json'))
print($t, {k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'])
PY
done
