#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_attention.py 64 256 2>&1 | tee gpurun_out/b_trace.log
timeout 1500 python -m pytest tests -m gpu -q -k "not api_scenarios and not clip_b8_t20" --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/b_tests.log
