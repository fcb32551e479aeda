#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'dec_tok_kernel' -s 16 -c 2 -o gpurun_out/r2_dec_tok -f python tools/trace_dec.py > gpurun_out/s_ncu.log 2>&1
tail -5 gpurun_out/s_ncu.log
