"""Roofline micro-benchmark of the bandwidth-bound kernels at shapes larger than L2 (see video_llava_seg_b200/kernel_bench.py).
    python tools/bench_kernels.py [out.json]        # all shapes, CUDA events
    python tools/bench_kernels.py --once            # every kernel launched a single time (for `ncu --set full`)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_llava_seg_b200 import kernel_bench

dev = torch.device("cuda:0")
peak = 6550.1
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
if "--once" in sys.argv:
    for name, fn, _ in kernel_bench.cases(dev, quick=False):
        fn()
    torch.cuda.synchronize()
else:
    res = kernel_bench.hbm_rooflines(dev, peak, quick=False)
    for r in res:
        print(r, flush=True)
    for a in sys.argv[1:]:
        if a.endswith(".json"):
            json.dump(dict(hbm_peak_GBps=peak, results=res), open(a, "w"), indent=1)
