#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "ffn_fused" --no-header -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/f_tests_ffn.log
timeout 300 python tools/timeline_frame.py > gpurun_out/f_timeline.txt 2>&1
grep -n "ffn_fused" gpurun_out/f_timeline.txt | head -3
timeout 900 python -m pytest tests -m gpu -q -k "seg_head or api_scenarios" --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/f_tests.log
