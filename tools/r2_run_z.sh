#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dwconv7" --no-header -p no:cacheprovider 2>&1 | tail -8
timeout 300 python tools/bench_dwconv.py 2>&1 | tail -5
