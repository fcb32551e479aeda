#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cc or fill or hole or connected" --no-header -p no:cacheprovider 2>&1 | tail -4
python tools/debug_cc.py 300 2>&1 | tail -3
timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "cc_label|fill_holes"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'cc_' -c 40 --csv --log-file gpurun_out/cc_launches.csv python tools/bench_kernels.py --once > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/cc_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:24]: print(r[4][:60], r[-1])
PY
touch video-llava-seg_b200/csrc/cc.cu
VLS_EXTRA_NVCC_FLAGS=-DCC_TRACE python -m video_llava_seg_b200.build > gpurun_out/y_build.log 2>&1; tail -1 gpurun_out/y_build.log
timeout 300 python tools/trace_cc.py > gpurun_out/y_trace_cc.log 2>&1; cat gpurun_out/y_trace_cc.log | head -8
