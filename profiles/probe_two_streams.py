"""Does overlapping two half-batches on two streams beat one batch on one stream?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from __graft_entry__ import load_package
pp = load_package()
n = 1 << 20
m = pp.Map()
fr = pp.synth_frames(m, n, 12)
whole = pp.DeviceFrames(fr); dpw = pp.DevicePlans(n, 12, diag=True, cars=False)
halves = [pp.DeviceFrames(fr.slice(0, n // 2)), pp.DeviceFrames(fr.slice(n // 2, n))]
dph = [pp.DevicePlans(n // 2, 12, diag=True, cars=False) for _ in range(2)]
s = [torch.cuda.Stream(), torch.cuda.Stream()]
def one():
    pp.plan_batch(m, whole, dpw)
def two():
    cur = torch.cuda.current_stream()
    for i in range(2):
        s[i].wait_stream(cur)
        pp.plan_batch(m, halves[i], dph[i], stream=s[i].cuda_stream)
    for i in range(2):
        cur.wait_stream(s[i])
for fn, name in ((one, "one stream"), (two, "two streams"), (one, "one stream"), (two, "two streams")):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1)/10:.3f} ms per 1M frames")
