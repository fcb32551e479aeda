#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "ffn_fused" --no-header -p no:cacheprovider 2>&1 | tail -20 | tee gpurun_out/d_tests_ffn.log
timeout 900 python -m pytest tests -m gpu -q -k "memory_attention or (propagation_matches and not b8_t20) or cuda_graph_steady" --no-header -p no:cacheprovider 2>&1 | tail -20 | tee gpurun_out/d_tests_parity.log
timeout 300 python tools/timeline_frame.py > gpurun_out/d_timeline.txt 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 2>&1 | tail -5 | tee gpurun_out/d_bench.log
