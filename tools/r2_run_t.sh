#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "mask_decoder_fused_paths or mid_fused or memory_attention_small" > gpurun_out/t_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/t_memcheck.log
timeout 900 python bench.py --steps 40 --warmup 5 2>gpurun_out/t_bench.err | tail -1 > gpurun_out/t_bench.json
cut -c1-600 gpurun_out/t_bench.json
timeout 300 python tools/timeline_frame.py > gpurun_out/t_timeline.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 4 --warmup 3 --repeats 1 --no-cpu-baseline --no-pixels > gpurun_out/t_ncu.log 2>&1
tail -2 gpurun_out/t_ncu.log | cut -c1-300
