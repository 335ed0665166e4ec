"""Host-side vs device-side time of pp_plan_batch (diagnostic)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pp = load_package()
v = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
cars = int(sys.argv[3]) if len(sys.argv) > 3 else 12
pp.set_kernel_variant(v)
m = pp.Map()
fr = pp.synth_frames(m, n, cars)
df = pp.DeviceFrames(fr); dp = pp.DevicePlans(n, cars, diag=True, cars=False)
for _ in range(3): pp.plan_batch(m, df, dp)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    for _ in range(5): pp.plan_batch(m, df, dp)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    if rep == 2:
        pp.set_phase_timing(True)
        for _ in range(5): pp.plan_batch(m, df, dp)
        ms, ch = pp.get_phase_ms()
        pp.set_phase_timing(False)
        print("phases ms/call: prep %.3f cars %.3f decide %.3f emit %.3f slow %.3f (chunks %d)" % (*[v/5 for v in ms], ch))
    print(f"variant {v} n {n} cars {cars}: host issue {1e3*(t1-t0)/5:.3f} ms/call, device {e0.elapsed_time(e1)/5:.3f} ms/call, wall {1e3*(t2-t0)/5:.3f} ms/call")
