#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/micro/mma_rate 2>&1 | tee gpurun_out/c_mma_rate.log
timeout 900 python -m pytest tests -m gpu -q -k "clip_b1_t20 or gate or interleaved" --no-header -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/c_tests.log
