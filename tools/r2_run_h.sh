#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "layer_tail or ffn_fused" --no-header -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/h_tests_tail.log
timeout 300 python tools/trace_ffn.py 2>&1 | tee gpurun_out/h_trace_ffn.log
timeout 600 python -m pytest tests -m gpu -q -k "api_scenarios" --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/h_tests_api.log
