"""Phase times of the pipeline at small batch sizes (what a rollout group's tick costs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pp = load_package()
m = pp.Map()
for n in (4096, 16384, 65536):
    fr = pp.synth_frames(m, n, 12)
    df = pp.DeviceFrames(fr); dp = pp.DevicePlans(n, 12, diag=True, cars=True)
    for _ in range(3): pp.plan_batch(m, df, dp)
    torch.cuda.synchronize()
    pp.set_pipes(1); pp.set_phase_timing(True)
    for _ in range(20): pp.plan_batch(m, df, dp)
    ms, ch = pp.get_phase_ms(); pp.set_phase_timing(False); pp.set_pipes(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): pp.plan_batch(m, df, dp)
    e1.record(); torch.cuda.synchronize()
    print(f"n={n}: prep %.1f cars %.1f decide %.1f emit %.1f us | whole call %.1f us" % (*[v / 20 * 1e3 for v in ms[:4]], e0.elapsed_time(e1) / 20 * 1e3))
