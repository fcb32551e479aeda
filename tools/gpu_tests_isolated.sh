#!/bin/bash
# Run groups of GPU tests in separate processes (a trapping kernel poisons its CUDA context) and
# keep every log under gpurun_out/.  Usage: tools/gpu_tests_isolated.sh [pytest -k expressions...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
i=0
for expr in "$@"; do
  i=$((i+1))
  echo "=== group $i: $expr"
  timeout 600 python -m pytest tests -m gpu -q -k "$expr" --no-header -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/tests_$i.log
done
