#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_dec.py > gpurun_out/r_trace.txt 2>&1
cat gpurun_out/r_trace.txt
