#!/bin/bash
mkdir -p gpurun_out
python tools/prof_kernels.py both > gpurun_out/e_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'ffn_fused_kernel|attn_fwd_kernel' -s 2 -c 4 -o gpurun_out/e_prof python tools/prof_kernels.py both > gpurun_out/e_ncu.log 2>&1
tail -5 gpurun_out/e_ncu.log
timeout 900 python -m pytest tests -m gpu -q -k "api_scenarios or clip_b8_t20" --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/e_tests.log
