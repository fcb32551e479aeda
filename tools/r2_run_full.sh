#!/bin/bash
# full GPU suite, smoke, default bench, kernel stress shapes, ncu captures of the two kernels changed in this session
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider ) > gpurun_out/f_tests.log 2>&1; tail -5 gpurun_out/f_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 600 python bench.py 2>gpurun_out/f_bench.err | tail -1 > gpurun_out/f_bench.json; tail -2 gpurun_out/f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], [ (r['kernel'], r['ms'], r['frac']) for r in d.get('roofline_hbm',[])])
PY
timeout 300 python tools/bench_kernels.py gpurun_out/f_kernels.json > gpurun_out/f_kernels.log 2>&1; cat gpurun_out/f_kernels.log
ncu --set full --clock-control none --import-source on -k regex:'cc_small_kernel' -c 1 -o gpurun_out/r2_cc_small -f python tools/bench_kernels.py --once > gpurun_out/f_ncu1.log 2>&1; tail -1 gpurun_out/f_ncu1.log
python tools/ncu_full_summary.py gpurun_out/r2_cc_small.ncu-rep > gpurun_out/r2_cc_small_summary.txt; tail -28 gpurun_out/r2_cc_small_summary.txt
