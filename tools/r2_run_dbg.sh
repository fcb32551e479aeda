#!/bin/bash
for flag in CC_DBG_CAS CC_DBG_NOCOMPRESS "CC_DBG_CAS -DCC_DBG_NOCOMPRESS"; do
  touch video-llava-seg_b200/csrc/cc.cu
  VLS_EXTRA_NVCC_FLAGS="-D$flag" python -m video_llava_seg_b200.build > /dev/null 2>&1
  echo "== $flag"; python tools/debug_cc.py 300 2>&1 | tail -3
done
