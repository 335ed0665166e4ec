"""Single-frame latency through the host entry point (the reference's use: one frame per 20 ms)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pp = load_package()
m = pp.Map()
for n in (1, 16, 256, 4096):
    fr = pp.synth_frames(m, n, 12)
    pl = pp.PlanBatch(n, 12, diag=True, cars=False)
    for _ in range(5): pp.plan_batch_host(m, fr, pl)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); pp.plan_batch_host(m, fr, pl); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(f"n={n:5d}: median {np.median(ts):8.1f} us  p99 {np.percentile(ts, 99):8.1f} us per call")
