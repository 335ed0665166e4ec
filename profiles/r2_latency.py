"""Device time of pp_plan_batch for small batches (CUDA events around back-to-back calls on
device-resident buffers): warp-per-frame kernel (variant 4, the default below 4096 frames)
against the thread-per-frame fused kernel (1) and the pipeline (2); plus the host entry point
(pinned buffers, copies included)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pp = load_package()
m = pp.Map()
for n in (1, 32, 256, 1024, 4095):
    fr = pp.synth_frames(m, n, 12, seed=5)
    df = pp.DeviceFrames(fr); dp = pp.DevicePlans(n, 12, diag=True, cars=False)
    row = []
    for v in (4, 1, 2):
        pp.set_kernel_variant(v)
        for _ in range(5): pp.plan_batch(m, df, dp)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): pp.plan_batch(m, df, dp)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
        row.append(best)
    pp.set_kernel_variant(0)
    hp = pp.PlanBatch(n, 12, diag=False, cars=False)
    for _ in range(3): pp.plan_batch_host(m, fr, hp)
    t0 = time.perf_counter()
    for _ in range(20): pp.plan_batch_host(m, fr, hp)
    host = (time.perf_counter() - t0) / 20 * 1e6
    print(f"n {n:5d}: device us/call  warp-per-frame {row[0]:8.1f}  thread-per-frame {row[1]:8.1f}  pipeline {row[2]:8.1f}   host entry point (default kernel, pageable buffers) {host:8.1f}")
