#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'mid_fused_kernel' -s 8 -c 1 -o gpurun_out/r2_mid_fused -f python tools/trace_mid.py > gpurun_out/v_ncu1.log 2>&1; tail -2 gpurun_out/v_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:'dec_img_kernel' -s 4 -c 1 -o gpurun_out/r2_dec_img -f python tools/trace_dec_img.py > gpurun_out/v_ncu2.log 2>&1; tail -2 gpurun_out/v_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:'dec_tok_kernel' -s 16 -c 2 -o gpurun_out/r2_dec_tok -f python tools/trace_dec_img.py > gpurun_out/v_ncu3.log 2>&1; tail -2 gpurun_out/v_ncu3.log
ncu --set full --clock-control none --import-source on -k regex:'up2_masks_tc_kernel' -s 2 -c 1 -o gpurun_out/r2_up2_tc -f python tools/trace_dec_img.py > gpurun_out/v_ncu4.log 2>&1; tail -2 gpurun_out/v_ncu4.log
