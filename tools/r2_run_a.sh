#!/bin/bash
# r2 GPU call A: new dv=64 attention (both value layouts), module + predictor parity, micro-bench, timeline, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "attention" --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/a_tests_attention.log
timeout 300 python tools/bench_attention.py 2>&1 | tee gpurun_out/a_bench_attention.log
timeout 900 python -m pytest tests -m gpu -q -k "memory_attention or propagation_matches or cuda_graph_steady or interleaved" --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/a_tests_parity.log
timeout 300 python tools/timeline_frame.py > gpurun_out/a_timeline.txt 2>&1
timeout 600 python bench.py --steps 40 --warmup 5 2>&1 | tail -5 | tee gpurun_out/a_bench.log
