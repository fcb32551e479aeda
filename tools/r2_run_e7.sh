#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or memory_attention or clip_b1 or steady" --no-header -p no:cacheprovider 2>&1 | tail -3
for v in 0 1 0 1; do echo "== gemm_wave_pick=$v"; VLS_TUNING="gemm_wave_pick=$v" timeout 600 python bench.py --no-cpu-baseline --no-pixels 2>gpurun_out/d_bench.err | tail -1 > gpurun_out/e_bench_$v.json
python - $v <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/e_bench_{sys.argv[1]}.json'))
print({k:d.get(k) for k in ('value','ms_per_step','windows_ms_per_step')}, d['e2e']['value'])
PY
done
