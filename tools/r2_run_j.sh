#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "attention_value_dim_64" --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/j_tests.log
timeout 600 python tools/bench_attention_x2.py 2>&1 | tee gpurun_out/j_attn_x2.log
