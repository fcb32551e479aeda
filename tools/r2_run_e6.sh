#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
