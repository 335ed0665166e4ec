"""Per-phase times of the pipeline at 64 cars per frame (BASELINE configs[4])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pp = load_package()
n, c = 1 << 19, 64
m = pp.Map()
fr = pp.synth_frames(m, n, c)
df = pp.DeviceFrames(fr); dp = pp.DevicePlans(n, c, diag=True, cars=False)
for _ in range(3): pp.plan_batch(m, df, dp)
torch.cuda.synchronize()
pp.set_phase_timing(True)
for _ in range(5): pp.plan_batch(m, df, dp)
ms, ch = pp.get_phase_ms()
print("C=64, %d frames: ms/call prep %.3f cars %.3f decide %.3f emit %.3f tail %.3f" % (n, *[v / 5 for v in ms]))
