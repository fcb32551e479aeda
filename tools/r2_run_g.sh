#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "layer_tail or ffn_fused" --no-header -p no:cacheprovider -s 2>&1 | tail -12 | tee gpurun_out/g_tests_tail.log
timeout 900 python -m pytest tests -m gpu -q -k "memory_attention or (propagation_matches and not b8_t20) or cuda_graph_steady or seg_head or api_scenarios" --no-header -p no:cacheprovider 2>&1 | tail -20 | tee gpurun_out/g_tests_parity.log
timeout 300 python tools/timeline_frame.py > gpurun_out/g_timeline.txt 2>&1
timeout 900 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-2500 | tee gpurun_out/g_bench.log
