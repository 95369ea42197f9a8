"""Where does the per-window latency at batch 1 go?    For single shots of the headline configuration: wall-clock
CUDA-event latency of every window decode next to the sum of its kernels' own durations (swd_set_profiling), so that launch
gaps and kernel execution can be told apart."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from slidingwindowdecoder_b200.sliding_window import SlidingWindowDecoder

bench.select_workload("c3_gdg")
plan = bench.build_plan()
swd = SlidingWindowDecoder(plan, decoder="gdg", device=0, streams=1, **bench.WL["kw"])
det, obs = swd.sample_device(400, seed=5)
decs = swd.unique_decoders()
rows = []
for rep in range(2):
    for i in range(400):
        for d in decs:
            d.kernel_times(); d.set_profiling(rep == 1)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plan.windows]
        swd.decode_device(det[i:i + 1].clone(), obs[i:i + 1].clone(), window_events=ev)
        torch.cuda.synchronize()
        lat = [a.elapsed_time(b) for a, b in ev]
        if rep == 1:
            kt = {}
            nl = 0
            for d in decs:
                for k, (t, ln) in d.kernel_times().items():
                    kt[k] = kt.get(k, 0.0) + t; nl += ln
            rows.append((sum(lat), max(lat), sum(kt.values()), nl, kt))
        else:
            rows.append((sum(lat), max(lat), None, None, None))
plain = np.array([r[1] for r in rows[:400]])
print(json.dumps({"max_window_latency_ms_unprofiled": {"p50": float(np.median(plain)), "p90": float(np.percentile(plain, 90)), "p99": float(np.percentile(plain, 99))}}))
prof = rows[400:]
worst = sorted(prof, key=lambda r: -r[0])[:5]
for w in worst:
    print(json.dumps({"shot_total_ms_with_event_pairs": round(w[0], 3), "worst_window_ms": round(w[1], 3), "sum_of_kernel_ms": round(w[2], 3), "launches": w[3],
                      "kernels": {k: round(v, 3) for k, v in w[4].items() if v}}))
