#!/bin/bash
# profile artefacts of the CC / dwconv work: stress test, event-timed stress shapes, phase trace, ncu --set full of both kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "stress or dwconv7" --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_kernels.py gpurun_out/r2_bandwidth_kernels_events.json > gpurun_out/p_kernels.log 2>&1; cat gpurun_out/p_kernels.log
ncu --set full --clock-control none --import-source on -k regex:'cc_small_kernel' -c 1 -o gpurun_out/r2_cc_small -f python tools/bench_kernels.py --once > gpurun_out/p_ncu1.log 2>&1; tail -1 gpurun_out/p_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:'dwconv7_ln_tma' -c 1 -o gpurun_out/r2_dwconv_tma -f python tools/bench_kernels.py --once > gpurun_out/p_ncu2.log 2>&1; tail -1 gpurun_out/p_ncu2.log
python tools/ncu_full_summary.py gpurun_out/r2_cc_small.ncu-rep > gpurun_out/r2_cc_small_summary.txt
python tools/ncu_full_summary.py gpurun_out/r2_dwconv_tma.ncu-rep > gpurun_out/r2_dwconv_tma_summary.txt
grep -E "duration|inst_executed.sum|issue_active|dram__bytes" gpurun_out/r2_cc_small_summary.txt gpurun_out/r2_dwconv_tma_summary.txt
touch video-llava-seg_b200/csrc/cc.cu
VLS_EXTRA_NVCC_FLAGS=-DCC_TRACE python -m video_llava_seg_b200.build > gpurun_out/p_build.log 2>&1
timeout 300 python tools/trace_cc.py > gpurun_out/r2_cc_phase_trace.txt 2>&1; head -20 gpurun_out/r2_cc_phase_trace.txt
