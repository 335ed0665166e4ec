"""End-to-end (host buffers) time of pp_plan_batch_host for the bench workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from __graft_entry__ import load_package
pp = load_package()
n = 1 << 20
m = pp.Map()
frames = pp.synth_frames(m, n, 12)
hf = pp.FrameBatch(n, 12); keep = []
for k, v in frames.arrays().items():
    t = torch.from_numpy(v).pin_memory(); keep.append(t); setattr(hf, k, t.numpy())
hp = pp.PlanBatch(n, 12, diag=True, cars=False)
for k in hp.fields:
    t = torch.from_numpy(getattr(hp, k)).pin_memory(); keep.append(t); setattr(hp, k, t.numpy())
pp.plan_batch_host(m, hf, hp); torch.cuda.synchronize()
best = 1e9
for _ in range(6):
    t0 = time.perf_counter(); pp.plan_batch_host(m, hf, hp); torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print(f"first={os.environ.get('PP_HOST_CHUNK_FIRST','-')} cap={os.environ.get('PP_HOST_CHUNK_CAP','-')}: best {best*1e3:.2f} ms -> {n/best/1e6:.1f} M frames/s")
