#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_image_encoder_gpu.py -q -x -s --no-header -p no:cacheprovider 2>&1 | tail -50 | tee gpurun_out/l_tests.log
