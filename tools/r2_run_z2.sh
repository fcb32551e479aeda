#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'dwconv7_ln_tma' -s 3 -c 1 -o gpurun_out/r2_dwconv_tma -f python tools/bench_dwconv.py > gpurun_out/z2_ncu.log 2>&1; tail -2 gpurun_out/z2_ncu.log
python tools/ncu_full_summary.py gpurun_out/r2_dwconv_tma.ncu-rep | tee gpurun_out/z2_summary.txt | tail -32
