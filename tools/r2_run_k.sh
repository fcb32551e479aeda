#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gemm_tn_kernel' -c 2 -o gpurun_out/r2_kproj -f python tools/profile_frame.py 1 > gpurun_out/k_ncu.log 2>&1; tail -2 gpurun_out/k_ncu.log
python tools/ncu_full_summary.py gpurun_out/r2_kproj.ncu-rep > gpurun_out/r2_kproj_summary.txt; head -34 gpurun_out/r2_kproj_summary.txt
ncu -i gpurun_out/r2_kproj.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[2:3]:
    d=dict(zip(h,r))
    for k in ['launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__occupancy_limit_blocks','launch__waves_per_multiprocessor','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__inst_executed_pipe_lsu.sum','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','smsp__inst_executed_op_global_ld.sum','smsp__inst_executed_op_global_st.sum','dram__bytes_write.sum','launch__shared_mem_per_block_dynamic']:
        print(k, d.get(k))
"
