#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 2>gpurun_out/w_bench2.err | tail -1 > gpurun_out/w_bench_2gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/w_bench_2gpu.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','windows_ms_per_step')}, d['e2e']['value'])
PY
tail -3 gpurun_out/w_bench2.err
