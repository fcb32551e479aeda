#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cc or fill or hole or connected" --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "cc_label|fill_holes"
touch video-llava-seg_b200/csrc/cc.cu
VLS_EXTRA_NVCC_FLAGS=-DCC_TRACE python -m video_llava_seg_b200.build > gpurun_out/y_build.log 2>&1
timeout 300 python tools/trace_cc.py 2>&1 | head -5
touch video-llava-seg_b200/csrc/cc.cu
VLS_EXTRA_NVCC_FLAGS="-DCC_F_DIRECT" python -m video_llava_seg_b200.build > gpurun_out/y_build.log 2>&1
echo "== CC_F_DIRECT"; timeout 300 python tools/bench_kernels.py 2>&1 | grep -E "cc_label N=5|fill_holes"
touch video-llava-seg_b200/csrc/cc.cu
VLS_EXTRA_NVCC_FLAGS="-DCC_F_DIRECT -DCC_TRACE" python -m video_llava_seg_b200.build > gpurun_out/y_build.log 2>&1
timeout 300 python tools/trace_cc.py 2>&1 | head -5
