#!/bin/bash
# phase trace of the small-image CC kernel (clock64 marks; library rebuilt on the box with -DCC_TRACE)
mkdir -p gpurun_out
touch video-llava-seg_b200/csrc/cc.cu
VLS_EXTRA_NVCC_FLAGS=-DCC_TRACE python -m video_llava_seg_b200.build > gpurun_out/y_build.log 2>&1; tail -1 gpurun_out/y_build.log
timeout 300 python tools/trace_cc.py > gpurun_out/y_trace_cc.log 2>&1; cat gpurun_out/y_trace_cc.log | head -40
