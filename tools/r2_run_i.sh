#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/i_tests.log
timeout 300 python tools/timeline_frame.py > gpurun_out/i_timeline.txt 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 2>&1 | tail -1 > gpurun_out/i_bench.json
cut -c1-1500 gpurun_out/i_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 4 --warmup 3 --repeats 1 --no-cpu-baseline > gpurun_out/i_ncu.log 2>&1
tail -2 gpurun_out/i_ncu.log | cut -c1-300
