#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_ffn.py 2>&1 | tail -18
timeout 900 python -m pytest tests -m gpu -q -x -s --no-header -p no:cacheprovider -k "ffn or layer_tail or memory_attention" 2>&1 | tail -9 | tee gpurun_out/r_tests1.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r_tests.log
timeout 300 python tools/timeline_frame.py > gpurun_out/r_timeline.txt 2>&1
grep -E "frame span" gpurun_out/r_timeline.txt
grep -E "ffn_fused" gpurun_out/r_timeline.txt | head -3
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-pixels 2>gpurun_out/r_bench.err | tail -1 > gpurun_out/r_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','windows_ms_per_step')}, d['e2e']['value'])
PY
