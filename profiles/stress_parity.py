"""One-off stress of GPU-vs-oracle parity over many frames and seeds (not part of the suite)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from __graft_entry__ import load_package
import checkers
from conftest import assert_plans_equal, plans_dict
pp = load_package()
m = pp.Map()
orc = checkers.Checker("oracle")
tot = 0
for seed, cars, rare, n in ((101, 12, 20, 400000), (102, 12, 300, 300000), (103, 64, 50, 100000), (104, 3, 500, 200000)):
    fb = pp.synth_frames(m, n, cars, seed=seed, rare_permille=rare, max_cars=cars)
    want = orc.plan(fb, threads=16)
    df = pp.DeviceFrames(fb); dp = pp.DevicePlans(n, cars, diag=True, cars=True)
    pp.plan_batch(m, df, dp); torch.cuda.synchronize()
    got = dp.to_host()
    assert_plans_equal(plans_dict(got), plans_dict(want), (1 << 21) - 1, bitwise_traj=False, what=f"seed {seed}: ")
    err = np.nanmax(np.abs(got.next_x - want.next_x))
    tot += n
    print(f"seed {seed} cars {cars} rare {rare}: {n} frames ok, max |dx| {err:.2e}")
print("stress parity ok:", tot, "frames")
