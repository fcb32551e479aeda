#!/bin/bash
mkdir -p gpurun_out
python tools/prof_kernels.py attn > gpurun_out/o_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'attn_x2_kernel' -s 2 -c 2 -o gpurun_out/r2_attn_x2 -f python tools/prof_kernels.py attn > gpurun_out/o_ncu.log 2>&1
tail -3 gpurun_out/o_ncu.log
