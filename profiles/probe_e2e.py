"""End-to-end probe of the host entry points (pinned buffers, 1,048,576 frames x 12 cars):
frames/s for whole rows and split rows under the chunk schedule of this process
(PP_HOST_CHUNK_FIRST / PP_HOST_CHUNK_CAP).  usage: python profiles/probe_e2e.py [steps]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pp = importlib.import_module("carnd-path-planning-project_b200")

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n, cars = 1 << 20, 12
pp.lib.pp_init()
torch.cuda.init()
m = pp.Map()
fb = pp.synth_frames(m, n, cars, seed=0x5EED)
keepalive = []


def pin(a):
    if os.environ.get("PIN", "pp") == "torch":  # torch's caching pinned allocator
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        keepalive.append(t)
        return t.numpy()
    return pp.pinned_like(np.ascontiguousarray(a))  # pp_host_alloc = cudaHostAlloc


hf = pp.FrameBatch(n, cars)
for k, v in fb.arrays().items():
    setattr(hf, k, pin(v))
hp = pp.PlanBatch(n, cars, diag=False, cars=False)
for k in hp.fields:
    setattr(hp, k, pin(getattr(hp, k)))
sp = pp.PlanBatch(n, cars, diag=False, cars=False)
sp.fields = [f for f in sp.fields if f not in ("next_x", "next_y")]
sp.next_x = sp.next_y = None
for k in sp.fields:
    setattr(sp, k, getattr(hp, k))
keep, tail = pp.PREV_KEEP, pp.PATH_LEN - pp.PREV_KEEP
tx, ty = pin(np.empty((n, tail))), pin(np.empty((n, tail)))
hx, hy = pin(hf.prev_x.copy()), pin(hf.prev_y.copy())


def run(call):
    call()
    torch.cuda.synchronize()
    best, tot = 1e9, 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        call()
        dt = time.perf_counter() - t0
        best, tot = min(best, dt), tot + dt
    return n / (tot / steps) / 1e6, n / best / 1e6


w = run(lambda: pp.plan_batch_host(m, hf, hp))
s = run(lambda: pp.plan_batch_host_split(m, hf, sp, hx, hy, tx, ty))
print(f"pin={os.environ.get('PIN', 'pp')} first={os.environ.get('PP_HOST_CHUNK_FIRST', 'dflt')} cap={os.environ.get('PP_HOST_CHUNK_CAP', 'dflt')}: "
      f"whole rows {w[0]:.1f} M/s (best {w[1]:.1f}), split rows {s[0]:.1f} M/s (best {s[1]:.1f})")
