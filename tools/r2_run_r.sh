#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "decoder" 2>&1 | tail -25 | tee gpurun_out/r_tests1.log
timeout 300 python tools/trace_dec.py > gpurun_out/r_trace.txt 2>&1
head -75 gpurun_out/r_trace.txt
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r_tests.log
timeout 300 python tools/timeline_frame.py > gpurun_out/r_timeline.txt 2>&1
grep -E "frame span" gpurun_out/r_timeline.txt
