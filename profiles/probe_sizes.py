"""pp_plan_batch device time for several batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pp = load_package()
m = pp.Map()
for n in (16384, 65536, 131072, 262144, 524288, 1 << 20, 1 << 21):
    fr = pp.synth_frames(m, n, 12)
    df = pp.DeviceFrames(fr); dp = pp.DevicePlans(n, 12, diag=True, cars=False)
    for _ in range(3): pp.plan_batch(m, df, dp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): pp.plan_batch(m, df, dp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"n={n:8d}: {ms:7.3f} ms  {n/ms/1e3:7.1f} M frames/s")
