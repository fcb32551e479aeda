#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cc or fill or hole or connected" --no-header -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "cc_label|fill_holes"
