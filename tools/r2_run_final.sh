#!/bin/bash
# final artefacts of the round: tests, smoke, bench lines (configs[1..4]), kernel stress shapes, warm timeline, ncu launch list
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider ) > gpurun_out/f_tests.log 2>&1; tail -5 gpurun_out/f_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
timeout 900 python bench.py 2>gpurun_out/f_bench.err | tail -1 > gpurun_out/r2_bench_final.json; tail -2 gpurun_out/f_bench.err
for c in 2 3 4; do timeout 900 python bench.py --workload "configs[$c]" --no-cpu-baseline --no-pixels 2>gpurun_out/f_c$c.err | tail -1 > gpurun_out/r2_bench_configs$c.json; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/f_ref.err | tail -1 > gpurun_out/r2_bench_reference_arm.json
python - <<'PY'
import json
for n in ('r2_bench_final','r2_bench_configs2','r2_bench_configs3','r2_bench_configs4','r2_bench_reference_arm'):
    try:
        d=json.load(open(f'gpurun_out/{n}.json'))
        print(n, d.get('value'), d.get('unit'), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))
    except Exception as e: print(n, 'ERR', e)
d=json.load(open('gpurun_out/r2_bench_final.json'))
print([ (r['kernel'], r['ms'], r['frac']) for r in d.get('roofline_hbm',[])]); print(d.get('cpu_baseline')); print(d.get('gpu_eager_baseline')); print(d.get('e2e_from_pixels')); print(d.get('clocks'))
PY
python tools/timeline_frame.py > gpurun_out/r2_timeline_final.txt 2>&1; grep "frame span" gpurun_out/r2_timeline_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2_launches_final_ncu_gpu_time.csv python tools/profile_frame.py 2 > gpurun_out/f_ncu.log 2>&1; tail -1 gpurun_out/f_ncu.log; wc -l gpurun_out/r2_launches_final_ncu_gpu_time.csv
