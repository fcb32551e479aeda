#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_ffn.py 2>&1 | tail -18
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "ffn or layer_tail or memory_attention or memory_encoder" 2>&1 | tail -4 | tee gpurun_out/r_tests1.log
