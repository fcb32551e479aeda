#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'mds1|mds2|dwconv7_ln_kernel|axpy_rows|im2col|rows_to_nchw|chlast' -c 14 -o gpurun_out/r2_memenc_small -f python tools/profile_frame.py 1 > gpurun_out/m_ncu.log 2>&1; tail -2 gpurun_out/m_ncu.log
python tools/ncu_full_summary.py gpurun_out/r2_memenc_small.ncu-rep > gpurun_out/r2_memenc_small_summary.txt
grep -E "^==|duration|inst_executed.sum|issue_active|warps_active|long_scoreboard|registers|grid_size|dram__bytes_read" gpurun_out/r2_memenc_small_summary.txt
