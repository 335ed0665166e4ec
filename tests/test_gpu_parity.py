"""Parity tests proper: the CUDA path (through the C ABI) against the CPU
checkers on the same inputs.

Bars (north_star): lane / waypoint / target-lane indices and every quantity
built only from + - * / sqrt bit-exact; trajectory x/y within 1e-9 relative /
1e-6 m absolute (they pass through atan2/sin/cos, where CUDA's libdevice and
glibc differ by an ulp or two).  /root/reference is never read here; the
reference's own code takes part through the prebuilt oracle/_ref/libppref.so
when it travelled with the snapshot, and always through tests/golden/*.npz.
"""
import ctypes as C

import numpy as np
import pytest

import checkers
from conftest import assert_plans_equal, plans_dict

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def gmap(pp, torch_cuda):
    m = pp.Map()
    assert pp.lib.pp_map_has_device(m.handle) == 1, pp.lib.pp_last_cuda_error().decode()
    return m


def gpu_plan(pp, torch, m, fb, cars=True):
    df = pp.DeviceFrames(fb)
    dp = pp.DevicePlans(fb.n, fb.max_cars, diag=True, cars=cars)
    before = pp.launch_count()
    pp.plan_batch(m, df, dp)
    torch.cuda.synchronize()
    assert pp.launch_count() > before  # our kernels ran (1 fused launch or 3 per pipeline chunk)
    return dp.to_host()


ALL_FLAGS = (1 << 21) - 1


def test_golden_frames(pp, torch_cuda, gmap, golden_frames):
    """GPU vs the committed outputs of the reference's own code."""
    fb, want, mask = golden_frames
    got = gpu_plan(pp, torch_cuda, gmap, fb)
    assert_plans_equal(plans_dict(got), want, mask, bitwise_traj=False, what="gpu vs golden: ")


@pytest.mark.parametrize("cars,n,seed,rare", [(12, 20000, 11, 100), (64, 3000, 12, 100),
                                              (12, 4000, 13, 1000), (0, 500, 14, 50),
                                              (1, 1000, 15, 200), (5, 2000, 16, 0)])
def test_gpu_vs_oracle(pp, torch_cuda, gmap, oracle, cars, n, seed, rare):
    fb = pp.synth_frames(gmap, n, cars, seed=seed, rare_permille=rare, max_cars=max(cars, 1))
    want = oracle.plan(fb, threads=8)
    got = gpu_plan(pp, torch_cuda, gmap, fb)
    assert_plans_equal(plans_dict(got), plans_dict(want), ALL_FLAGS, bitwise_traj=False)


def test_gpu_vs_reference_binary(pp, torch_cuda, gmap):
    """Directly against the reference's compiled classes, when the prebuilt
    harness travelled with the snapshot."""
    if not checkers.available("ref"):
        pytest.skip("oracle/_ref/libppref.so not in the snapshot")
    ref = checkers.Checker("ref")
    fb = pp.synth_frames(gmap, 8000, 12, seed=21, rare_permille=150)
    want = ref.plan(fb, threads=8, want_flags=False)
    got = gpu_plan(pp, torch_cuda, gmap, fb)
    # flags: the multi-threaded reference run keeps its log sites off, so compare the printf-visible bits only
    mask = sum(pp.FLAG[k] for k in ("EGO_MATCH_FAIL", "CAR_DROPPED", "COLLISION", "SPLINE_INPUT_ERR"))
    assert_plans_equal(plans_dict(got), plans_dict(want), mask, bitwise_traj=False)


def test_fused_and_pipeline_variants_are_bitwise_identical(pp, torch_cuda, gmap):
    """The single fused kernel (variant 1), the strided pipeline (variant 2) and the tiled
    pipeline (variant 3, the default for large batches) run the same __device__ code with
    different thread mappings: every output bit must agree."""
    fb = pp.synth_frames(gmap, 300000, 12, seed=61, rare_permille=100)  # > 1 pipeline chunk
    try:
        pp.set_kernel_variant(1)
        a = gpu_plan(pp, torch_cuda, gmap, fb)
        pp.set_kernel_variant(2)
        b = gpu_plan(pp, torch_cuda, gmap, fb)
        pp.set_kernel_variant(3)
        c = gpu_plan(pp, torch_cuda, gmap, fb)
    finally:
        pp.set_kernel_variant(0)
    for k in a.fields:
        assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), k
        assert np.array_equal(getattr(a, k), getattr(c, k), equal_nan=True), k


@pytest.mark.parametrize("cars,n", [(0, 4500), (1, 5003), (2, 4097), (5, 4999), (12, 6001),
                                    (20, 4100), (33, 4101), (63, 4200), (64, 4203)])
def test_pipeline_tile_geometries(pp, torch_cuda, gmap, oracle, cars, n):
    """Every tile geometry of the tiled pipeline (frames per warp tile, lanes per frame in the
    reduction, odd car counts that cannot be bulk-copied, ragged last tiles) and the strided
    pipeline at the same sizes (max_cars = 1 once divided by zero there): bit for bit against
    the fused kernel, and against the oracle."""
    rng = np.random.default_rng(cars)
    fb = pp.synth_frames(gmap, n, cars, seed=300 + cars, rare_permille=100, max_cars=max(cars, 1))
    if cars > 1:
        fb.n_cars[::3] = rng.integers(0, cars + 1, len(fb.n_cars[::3]))
    try:
        pp.set_kernel_variant(1)
        a = gpu_plan(pp, torch_cuda, gmap, fb)
        pp.set_kernel_variant(2)
        b = gpu_plan(pp, torch_cuda, gmap, fb)
        pp.set_kernel_variant(3)
        c = gpu_plan(pp, torch_cuda, gmap, fb)
    finally:
        pp.set_kernel_variant(0)
    for k in a.fields:
        if k.startswith("car_"):  # slots at and beyond n_cars are not written
            live = np.arange(fb.max_cars)[None, :] < np.minimum(fb.n_cars, fb.max_cars)[:, None]
            x, y, z = getattr(a, k)[live], getattr(b, k)[live], getattr(c, k)[live]
        else:
            x, y, z = getattr(a, k), getattr(b, k), getattr(c, k)
        assert np.array_equal(x, y, equal_nan=True), ("strided", k)
        assert np.array_equal(x, z, equal_nan=True), ("tiled", k)
    sub = fb.slice(0, 1500)
    want = oracle.plan(sub, threads=8)
    got = {k: getattr(c, k)[:1500] for k in c.fields}
    assert_plans_equal(got, plans_dict(want), ALL_FLAGS, bitwise_traj=False)


@pytest.mark.parametrize("cars,n", [(0, 70), (1, 333), (12, 2500), (40, 300), (64, 257)])
def test_warp_per_frame_kernel_is_bitwise_the_fused_kernel(pp, torch_cuda, gmap, cars, n):
    """Variant 4 (one warp per frame: the scan, the cars and the reductions spread over the
    lanes; the default below 4096 frames) against variant 1 (one thread per frame): every bit,
    with ragged car counts, more cars than lanes, and rare frames."""
    rng = np.random.default_rng(cars + 7)
    fb = pp.synth_frames(gmap, n, cars, seed=400 + cars, rare_permille=200, max_cars=max(cars, 1))
    if cars > 1:
        fb.n_cars[::2] = rng.integers(0, cars + 1, len(fb.n_cars[::2]))
    try:
        pp.set_kernel_variant(1)
        a = gpu_plan(pp, torch_cuda, gmap, fb)
        pp.set_kernel_variant(4)
        b = gpu_plan(pp, torch_cuda, gmap, fb)
    finally:
        pp.set_kernel_variant(0)
    auto = gpu_plan(pp, torch_cuda, gmap, fb)
    live = np.arange(fb.max_cars)[None, :] < np.minimum(fb.n_cars, fb.max_cars)[:, None]
    for k in a.fields:
        x, y, z = getattr(a, k), getattr(b, k), getattr(auto, k)
        if k.startswith("car_"):
            x, y, z = x[live], y[live], z[live]
        assert np.array_equal(x, y, equal_nan=True), k
        assert np.array_equal(x, z, equal_nan=True), ("auto", k)


def test_unaligned_buffers_take_the_plain_copies(pp, torch_cuda, gmap):
    """Frame and plan arrays that start 8 bytes off a 16-byte boundary cannot be moved by TMA
    bulk copies: the kernels fall back to ordinary loads / stores, with identical results."""
    torch = torch_cuda
    n = 5000
    fb = pp.synth_frames(gmap, n, 12, seed=71, rare_permille=100)
    want = gpu_plan(pp, torch, gmap, fb, cars=False)
    df = pp.DeviceFrames(fb)
    for k, v in list(df.t.items()):  # same values, every array shifted by one element
        big = torch.empty(v.numel() + 4, dtype=v.dtype, device=v.device)
        off = 1 if v.element_size() == 8 else 2
        big[off:off + v.numel()] = v.reshape(-1)
        df.t[k] = big[off:off + v.numel()].view(v.shape)
        assert df.t[k].data_ptr() % 16 == 8
    dp = pp.DevicePlans(n, 12, diag=True, cars=False)
    for k, v in list(dp.t.items()):
        big = torch.zeros(v.numel() + 4, dtype=v.dtype, device=v.device)
        off = 1 if v.element_size() == 8 else 2
        dp.t[k] = big[off:off + v.numel()].view(v.shape)
    try:
        pp.set_kernel_variant(3)
        pp.plan_batch(gmap, df, dp)
        torch.cuda.synchronize()
    finally:
        pp.set_kernel_variant(0)
    got = dp.to_host()
    for k in want.fields:
        assert np.array_equal(getattr(got, k), getattr(want, k), equal_nan=True), k


def test_ragged_and_edge_inputs(pp, torch_cuda, gmap, oracle):
    """Ragged n_cars per frame, ids in random order, empty batch, single frame."""
    rng = np.random.default_rng(3)
    fb = pp.synth_frames(gmap, 3000, 24, seed=31, rare_permille=100)
    fb.n_cars[:] = rng.integers(0, 25, fb.n)
    for f in range(0, fb.n, 3):  # shuffle car order within a frame (ids stay distinct)
        k = fb.n_cars[f]
        perm = rng.permutation(k)
        for name in ("car_id", "car_x", "car_y", "car_vx", "car_vy"):
            a = getattr(fb, name)
            a[f, :k] = a[f, :k][perm]
    want = oracle.plan(fb, threads=8)
    got = gpu_plan(pp, torch_cuda, gmap, fb)
    assert_plans_equal(plans_dict(got), plans_dict(want), ALL_FLAGS, bitwise_traj=False)
    # the order of the cars must not matter at all on the GPU (bitwise)
    fb2 = fb.slice(0, fb.n)
    for f in range(fb2.n):
        k = fb2.n_cars[f]
        for name in ("car_id", "car_x", "car_y", "car_vx", "car_vy"):
            a = getattr(fb2, name)
            a[f, :k] = a[f, :k][::-1]
    got2 = gpu_plan(pp, torch_cuda, gmap, fb2, cars=False)
    for k in ("next_x", "next_y", "target_lane", "flags", "target_speed", "target_time"):
        assert np.array_equal(getattr(got, k), getattr(got2, k), equal_nan=True), k
    one = fb.slice(5, 6)
    g1 = gpu_plan(pp, torch_cuda, gmap, one)
    assert np.array_equal(g1.next_x[0], got.next_x[5], equal_nan=True)
    # n = 0 is a no-op
    df = pp.DeviceFrames(one)
    dp = pp.DevicePlans(1, one.max_cars)
    pp.plan_batch(gmap, df, dp, n=0)
    torch_cuda.cuda.synchronize()
    assert int(dp.t["n_points"][0]) == 0


@pytest.mark.parametrize("cold", ["few", "some", "most"])
def test_host_entry_point_equals_device_entry_point(pp, torch_cuda, gmap, cold):
    """pp_plan_batch_host (copies + chunk pipeline inside) == pp_plan_batch, with few, some and
    mostly cold-start frames; a second call over stale host buffers gives the same."""
    n = 150000  # > 2 chunks of 65536, ragged tail
    fb = pp.synth_frames(gmap, n, 12, seed=41)
    rng = np.random.default_rng(41)
    if cold == "some":
        fb.prev_n[rng.random(n) < 0.1] = 3
    elif cold == "most":
        fb.prev_n[rng.random(n) < 0.6] = 0
        fb.prev_n[:40000] = 47  # ... but not in the first chunk
    dev = gpu_plan(pp, torch_cuda, gmap, fb, cars=True)
    host = pp.plan_batch_host(gmap, fb)
    host.next_x[:] = -7.0  # stale contents must not survive a second call either
    host.next_y[:] = -7.0
    host = pp.plan_batch_host(gmap, fb, plans=host)
    for k in host.fields:
        assert np.array_equal(getattr(host, k), getattr(dev, k), equal_nan=True), k


@pytest.mark.parametrize("cold", ["none", "few", "some", "most"])
def test_split_rows_host_entry_point(pp, torch_cuda, gmap, cold):
    """pp_plan_batch_host_split: head ++ tail == the whole rows of pp_plan_batch_host, bit for
    bit; head rows of frames that kept their previous points are not touched, so passing the
    frames' own prev_x / prev_y as the heads leaves the first points of every trajectory there
    (result_points = prev_trajectory, src/main.cpp:578)."""
    n = 150000
    fb = pp.synth_frames(gmap, n, 12, seed=43, rare_permille=0 if cold == "none" else 20)
    rng = np.random.default_rng(43)
    if cold == "some":
        fb.prev_n[rng.random(n) < 0.1] = 3
    elif cold == "most":
        fb.prev_n[rng.random(n) < 0.6] = 0
        fb.prev_n[:40000] = 47
    elif cold == "none":
        fb.prev_n[:] = np.maximum(fb.prev_n, pp.PREV_KEEP)
    whole = pp.plan_batch_host(gmap, fb, diag=False, cars=False)
    keep, tail_len = pp.PREV_KEEP, pp.PATH_LEN - pp.PREV_KEEP
    sp = pp.PlanBatch(n, fb.max_cars, diag=False, cars=False)
    sp.next_x = sp.next_y = None
    sp.fields = [f for f in sp.fields if f not in ("next_x", "next_y")]
    sentinel = -12345.0
    head_x = np.full((n, keep), sentinel)
    head_y = np.full((n, keep), sentinel)
    tail_x = np.full((n, tail_len), sentinel)
    tail_y = np.full((n, tail_len), sentinel)
    pp.plan_batch_host_split(gmap, fb, sp, head_x, head_y, tail_x, tail_y)
    for k in sp.fields:
        assert np.array_equal(getattr(sp, k), getattr(whole, k)), k
    kept = fb.prev_n >= keep
    assert np.array_equal(tail_x, whole.next_x[:, keep:], equal_nan=True)
    assert np.array_equal(tail_y, whole.next_y[:, keep:], equal_nan=True)
    assert np.all(head_x[kept] == sentinel) and np.all(head_y[kept] == sentinel)  # not touched
    assert np.array_equal(head_x[~kept], whole.next_x[~kept, :keep], equal_nan=True)
    assert np.array_equal(head_y[~kept], whole.next_y[~kept, :keep], equal_nan=True)
    # the kept points ARE the caller's previous points
    full = whole.n_points == pp.PATH_LEN
    assert np.array_equal(whole.next_x[kept & full, :keep], fb.prev_x[kept & full], equal_nan=True)
    assert np.array_equal(whole.next_y[kept & full, :keep], fb.prev_y[kept & full], equal_nan=True)
    # heads aliased onto the inputs: on return they hold the first points of every trajectory
    fb2 = fb.slice(0, n)
    fb2.prev_x, fb2.prev_y = fb.prev_x.copy(), fb.prev_y.copy()
    pp.plan_batch_host_split(gmap, fb2, sp, fb2.prev_x, fb2.prev_y, tail_x, tail_y)
    assert np.array_equal(fb2.prev_x[full], whole.next_x[full, :keep], equal_nan=True)
    assert np.array_equal(fb2.prev_y[full], whole.next_y[full, :keep], equal_nan=True)
    assert np.array_equal(tail_x, whole.next_x[:, keep:], equal_nan=True)
    # whole rows asked for as well: refused
    bad = pp.PlanBatch(8, fb.max_cars, diag=False, cars=False)
    with pytest.raises(pp.PPError):
        pp.plan_batch_host_split(gmap, fb.slice(0, 8), bad, head_x, head_y, tail_x, tail_y)


def test_host_entry_points_with_page_locked_buffers(pp, torch_cuda, gmap):
    """Page-locked caller buffers (pp_host_alloc): the per-frame scalars are then read and
    written in place by a kernel instead of the copy engines.  Same bytes as with pageable
    buffers, diagnostics and per-car outputs included."""
    n = 140001
    fb = pp.synth_frames(gmap, n, 12, seed=47)
    fb.prev_n[::7] = 2
    want = pp.plan_batch_host(gmap, fb)  # pageable numpy arrays: copies only
    pf = pp.FrameBatch(n, fb.max_cars)
    for k, v in fb.arrays().items():
        setattr(pf, k, pp.pinned_like(v))
    # one page-locked array that is NOT aligned to its elements: it takes the copy engine
    raw = pp.pinned_empty(n * 8 + 8, np.uint8)
    odd = raw[4:4 + n * 8].view(np.float64)
    odd[:] = fb.ego_y
    assert odd.ctypes.data % 8 == 4
    pf.ego_y = odd
    got = pp.PlanBatch(n, fb.max_cars)
    for k in got.fields:
        setattr(got, k, pp.pinned_like(getattr(got, k)))
    pp.plan_batch_host(gmap, pf, got)
    for k in got.fields:
        assert np.array_equal(getattr(got, k), getattr(want, k), equal_nan=True), k
    keep = pp.PREV_KEEP
    sp = pp.PlanBatch(n, fb.max_cars, cars=False)
    sp.fields = [f for f in sp.fields if f not in ("next_x", "next_y")]
    sp.next_x = sp.next_y = None
    for k in sp.fields:
        setattr(sp, k, pp.pinned_like(getattr(sp, k)))
    tx, ty = pp.pinned_empty((n, pp.PATH_LEN - keep)), pp.pinned_empty((n, pp.PATH_LEN - keep))
    pp.plan_batch_host_split(gmap, pf, sp, pf.prev_x, pf.prev_y, tx, ty)
    for k in sp.fields:
        assert np.array_equal(getattr(sp, k), getattr(want, k), equal_nan=True), k
    full = want.n_points == pp.PATH_LEN
    assert np.array_equal(tx, want.next_x[:, keep:], equal_nan=True)
    assert np.array_equal(ty, want.next_y[:, keep:], equal_nan=True)
    assert np.array_equal(pf.prev_x[full], want.next_x[full, :keep], equal_nan=True)
    assert np.array_equal(pf.prev_y[full], want.next_y[full, :keep], equal_nan=True)


def test_properties_at_full_size(pp, torch_cuda, gmap, oracle):
    """BASELINE config 2 size (1,048,576 frames, 12 cars): properties that do
    not need the CPU to plan a million frames, plus a 1/64 sample that does."""
    n = 1 << 20
    fb = pp.synth_frames(gmap, n, 12, seed=0x5EED)
    df = pp.DeviceFrames(fb)
    dp = pp.DevicePlans(n, 12, diag=True, cars=False)
    pp.plan_batch(gmap, df, dp)
    torch_cuda.cuda.synchronize()
    a = dp.to_host()
    # determinism: a second launch is bit-identical
    pp.plan_batch(gmap, df, dp)
    torch_cuda.cuda.synchronize()
    b = dp.to_host()
    for k in a.fields:
        assert np.array_equal(getattr(a, k), getattr(b, k), equal_nan=True), k
    # shard invariance: planning two halves separately gives the same plans and the same stats
    stats_whole = pp.stats_batch(dp).cpu().numpy()
    halves = []
    for lo, hi in ((0, n // 2), (n // 2, n)):
        part = fb.slice(lo, hi)
        dfp = pp.DeviceFrames(part)
        dpp = pp.DevicePlans(part.n, 12, diag=True, cars=False)
        pp.plan_batch(gmap, dfp, dpp)
        halves.append((dpp.to_host(), pp.stats_batch(dpp).cpu().numpy()))
    assert np.array_equal(np.concatenate([h[0].next_x for h in halves]), a.next_x, equal_nan=True)
    assert np.array_equal(np.concatenate([h[0].target_lane for h in halves]), a.target_lane)
    assert np.array_equal(halves[0][1] + halves[1][1], stats_whole)
    # invariants of the algorithm
    assert ((a.n_points >= 10) | (a.flags & pp.FLAG["COLD_START"]).astype(bool)).all()
    assert (a.n_points <= 50).all() and ((a.target_lane >= 0) & (a.target_lane <= 2)).all()
    assert (np.abs(a.target_lane - a.ego_lane) <= 1).all()  # adjacent-lane rule, src/main.cpp:473-479
    has_prev = fb.prev_n >= 10
    assert np.array_equal(a.next_x[has_prev, :10], fb.prev_x[has_prev])  # kept points are copied verbatim
    step = np.hypot(np.diff(a.next_x, axis=1), np.diff(a.next_y, axis=1))[:, 10:]
    ok_rows = ~(a.flags & (pp.FLAG["FALLBACK"] | pp.FLAG["EGO_MATCH_FAIL"])).astype(bool)
    assert np.nanmax(step[ok_rows]) <= 30.0 / 50  # 0.02 s steps: never faster than 30 m/s
    # stats kernel against numpy
    assert stats_whole[0] == n and stats_whole[1] == a.n_points.sum()
    assert list(stats_whole[2:5]) == list(np.bincount(a.target_lane, minlength=3))
    assert list(stats_whole[5:8]) == list(np.bincount(a.ego_lane, minlength=3))
    assert stats_whole[8] == (a.target_lane != a.ego_lane).sum()
    for bit in range(pp.NUM_FLAGS):
        assert stats_whole[9 + bit] == ((a.flags >> bit) & 1).sum(), pp.FLAG_NAMES[bit]
    # a strided 1/64 sample against the oracle
    idx = np.arange(0, n, 64)
    sample = pp.FrameBatch(len(idx), 12)
    for k, v in fb.arrays().items():
        setattr(sample, k, np.ascontiguousarray(v[idx]))
    want = oracle.plan(sample, threads=8, cars=False)
    got = {k: getattr(a, k)[idx] for k in a.fields}
    assert_plans_equal(got, plans_dict(want), ALL_FLAGS, bitwise_traj=False)


def test_closed_loop_teacher_forced(pp, torch_cuda, gmap, oracle, abi):
    """BASELINE config 3 in miniature: 256 egos x 120 ticks.  Each tick the GPU
    plans all egos; the simulator model consumes 3 points, the rest becomes the
    previous path, target_lane is carried, traffic advances at constant speed.
    The oracle is teacher-forced on the GPU's own frames each tick: integers
    bit-exact, trajectories within tolerance."""
    E, T, consumed = 256, 120, 3
    cur = pp.synth_frames(gmap, E, 12, seed=51, rare_permille=0)
    cur.prev_n[:] = 0
    cur.target_lane_in[:] = 1
    changes = 0
    for t in range(T):
        got = gpu_plan(pp, torch_cuda, gmap, cur, cars=False)
        want = oracle.plan(cur, threads=8, cars=False)
        assert_plans_equal(plans_dict(got), plans_dict(want), ALL_FLAGS, bitwise_traj=False,
                           what=f"tick {t}: ")
        changes += int((got.target_lane != cur.target_lane_in).sum())
        nxt = cur.slice(0, E)
        nxt.prev_n[:] = got.n_points - consumed
        nxt.prev_x[:] = got.next_x[:, consumed:consumed + 10]
        nxt.prev_y[:] = got.next_y[:, consumed:consumed + 10]
        nxt.ego_x[:], nxt.ego_y[:] = got.next_x[:, consumed - 1], got.next_y[:, consumed - 1]
        nxt.target_lane_in[:] = got.target_lane
        nxt.car_x += nxt.car_vx * 0.02 * consumed
        nxt.car_y += nxt.car_vy * 0.02 * consumed
        cur = nxt
    assert changes > 0  # lane changes did happen somewhere in the rollout


@pytest.mark.parametrize("kind,n_wp", [("circle", 40), ("ellipse", 400), ("circle", 7)])
def test_other_maps(pp, torch_cuda, oracle, kind, n_wp):
    """Tracks other than highway_map.csv: a small loop (every walk crosses the wrap-around seam,
    the table is shorter than the staging pad), a long one (table past the default shared-memory
    size), a degenerate 7-point loop (walks lap the whole table: the reference's unsigned index
    arithmetic, src/main.cpp:134-137)."""
    th = np.linspace(0.0, 2 * np.pi, n_wp, endpoint=False)
    if kind == "circle":
        r = 40.0 * n_wp / (2 * np.pi)           # ~40 m segments
        wx, wy = 1000 + r * np.cos(-th), 2000 + r * np.sin(-th)   # clockwise, like the highway
    else:
        wx, wy = 3000 + 2400 * np.cos(-th), 1000 + 1100 * np.sin(-th)
    chk = checkers.Checker("oracle")
    chk.map_from_points(wx, wy)
    m = pp.Map(points=(wx, wy))
    assert np.array_equal(m.table(), chk.map_table())
    for seed, cars, rare in ((1, 12, 100), (2, 64, 0)):
        fb = pp.synth_frames(m, 2500, cars, seed=seed, rare_permille=rare, max_cars=cars)
        want = chk.plan(fb, threads=8)
        got = gpu_plan(pp, torch_cuda, m, fb)
        assert_plans_equal(plans_dict(got), plans_dict(want), ALL_FLAGS, bitwise_traj=False,
                           what=f"{kind}/{n_wp}/seed {seed}: ")


@pytest.mark.parametrize("n", [700, 300000])
def test_plan_stats_batch_equals_the_two_calls(pp, torch_cuda, gmap, n):
    """pp_plan_stats_batch (statistics taken per chunk on the internal streams) against
    pp_plan_batch followed by pp_stats_batch: same plans, same statistics vector."""
    fb = pp.synth_frames(gmap, n, 12, seed=31, rare_permille=80)
    df = pp.DeviceFrames(fb)
    a = pp.DevicePlans(n, 12, diag=True, cars=False)
    b = pp.DevicePlans(n, 12, diag=True, cars=False)
    pp.plan_batch(gmap, df, a)
    want = pp.stats_batch(a).cpu().numpy()
    got = pp.plan_stats_batch(gmap, df, b).cpu().numpy()
    assert np.array_equal(got, want) and got[0] == n
    ha, hb = a.to_host(), b.to_host()
    for k in ha.fields:
        assert np.array_equal(getattr(ha, k), getattr(hb, k), equal_nan=True), k


def test_standstill_and_constant_speed_frames(pp, torch_cuda, gmap, oracle):
    """Operands the emission kernel's lean arithmetic hands to the complete path or treats
    specially: a car standing at keeping distance behind a stopped car (target speed 0 -> zero
    arc step -> 0/0 chord ratio: the reference emits ONE NaN point and stops), and cruising
    exactly at the target speed (zero numerators in the speed ramp).  Same bars as everywhere
    else, NaN pattern included."""
    cfg = pp.default_config()
    fb = pp.synth_frames(gmap, 3000, 12, seed=77, rare_permille=0)
    yaw = np.deg2rad(fb.ego_yaw_deg[:1000])
    fb.ego_speed_mph[:1000] = 0.0
    fb.prev_n[:1000] = 0
    fb.n_cars[:1000] = 1  # one stopped car inside the keep-distance leeway (src/main.cpp:1130-1136)
    fb.car_x[:1000, 0] = fb.ego_x[:1000] + 14.7 * np.cos(yaw)
    fb.car_y[:1000, 0] = fb.ego_y[:1000] + 14.7 * np.sin(yaw)
    fb.car_vx[:1000, 0] = 0.0
    fb.car_vy[:1000, 0] = 0.0
    fb.n_cars[1000:2000] = 0      # free road, no previous path, already at the speed limit
    fb.prev_n[1000:2000] = 0
    fb.ego_speed_mph[1000:2000] = cfg.max_speed / 0.44704
    assert fb.ego_speed_mph[1000] * 0.44704 == cfg.max_speed
    want = oracle.plan(fb, threads=8)
    # the constructed cases do occur in the reference's own arithmetic
    assert ((want.target_speed[:1000] == 0) & (want.n_points[:1000] == 1)).sum() > 800
    assert np.isnan(want.next_x[:1000, 0]).sum() > 800
    assert (want.target_speed[1000:2000] == cfg.max_speed).all()
    got = gpu_plan(pp, torch_cuda, gmap, fb)
    assert_plans_equal(plans_dict(got), plans_dict(want), ALL_FLAGS, bitwise_traj=False)
    # fused single-kernel path: bitwise the same as the pipeline here as well
    assert pp.lib.pp_set_kernel_variant(1) == 0
    try:
        fused = gpu_plan(pp, torch_cuda, gmap, fb)
    finally:
        pp.lib.pp_set_kernel_variant(0)
    for k in ("next_x", "next_y", "n_points", "flags", "target_lane"):
        assert np.array_equal(getattr(got, k), getattr(fused, k), equal_nan=True), k


def test_scattered_ego_positions(pp, torch_cuda, gmap, oracle):
    """Closest-waypoint search and everything behind it for ego positions that the synthetic
    generator never produces: scattered kilometres beyond the track, exactly on waypoints and
    midway between two (ties in distance).  Pipeline against the oracle and against the fused
    kernel, bit for bit; plus non-finite poses, which must not hang anything."""
    rng = np.random.default_rng(5)
    n = 60000
    fb = pp.synth_frames(gmap, n, 4, seed=91, rare_permille=0)
    fb.prev_n[:] = 0  # the ego position is the telemetry position
    tab = gmap.table()  # columns 0, 1: the waypoints
    wx, wy = tab[:, 0].copy(), tab[:, 1].copy()
    k = n // 6
    fb.ego_x[:k] = rng.uniform(wx.min() - 3000, wx.max() + 3000, k)       # anywhere, mostly far
    fb.ego_y[:k] = rng.uniform(wy.min() - 3000, wy.max() + 3000, k)
    i = rng.integers(0, len(wx), k)
    fb.ego_x[k:2 * k], fb.ego_y[k:2 * k] = wx[i], wy[i]                   # on a waypoint
    j = (i + 1) % len(wx)
    fb.ego_x[2 * k:3 * k] = (wx[i] + wx[j]) / 2                             # midway: near-ties
    fb.ego_y[2 * k:3 * k] = (wy[i] + wy[j]) / 2
    # poisoned poses: the reference never returns from get_lane_pos for a NaN pose
    # (src/main.cpp:282-325); this implementation and its oracle leave that loop, so the frame
    # costs nothing worse than NaN points — above all the kernel must come back
    bad = slice(3 * k, 3 * k + 6)
    fb.ego_x[bad] = [np.nan, 0.0, np.nan, np.inf, -np.inf, 1e200]
    fb.ego_y[bad] = [0.0, np.nan, np.nan, 0.0, np.inf, 1e200]
    want = oracle.plan(fb, threads=8)
    got = gpu_plan(pp, torch_cuda, gmap, fb)
    assert np.array_equal(got.ref_wp, want.ref_wp), np.argwhere(got.ref_wp != want.ref_wp)[:5]
    assert np.array_equal(got.ego_lane, want.ego_lane) and np.array_equal(got.target_lane, want.target_lane)
    assert np.array_equal(got.n_points, want.n_points)
    on_track = slice(k, 3 * k)  # full bars where the plan is well conditioned
    assert_plans_equal({a: v[on_track] for a, v in plans_dict(got).items()},
                       {a: v[on_track] for a, v in plans_dict(want).items()}, ALL_FLAGS,
                       bitwise_traj=False)
    try:
        pp.set_kernel_variant(1)
        fused = gpu_plan(pp, torch_cuda, gmap, fb)
    finally:
        pp.set_kernel_variant(0)
    for name in got.fields:
        assert np.array_equal(getattr(got, name), getattr(fused, name), equal_nan=True), name


def test_poisoned_fields_terminate_and_match(pp, torch_cuda, gmap, oracle):
    """NaN / infinite / huge / denormal values in every input field (car position and velocity,
    yaw, speed, previous path): the kernels return and every integer output, NaN pattern included,
    equals the oracle's — the select-based forms must treat NaN like the reference's branches."""
    n = 6400
    fb = pp.synth_frames(gmap, n, 4, seed=92, rare_permille=0)
    vals = [np.nan, np.inf, -np.inf, 1e200, 1e-310]
    fields = ["car_x", "car_y", "car_vx", "car_vy", "ego_yaw_deg", "ego_speed_mph", "prev_x", "prev_y"]
    for f in range(n):
        a = getattr(fb, fields[(f // 5) % 8])
        if a.ndim == 2:
            a[f, (f // 40) % a.shape[1]] = vals[f % 5]
        else:
            a[f] = vals[f % 5]
    fb.prev_n[::3] = 0  # cold starts take the telemetry speed: 1e200 mph makes get_lane_pos's s
    #                     so large that subtracting a segment no longer changes it
    want = oracle.plan(fb, threads=8)
    got = gpu_plan(pp, torch_cuda, gmap, fb)
    for name in ("ref_wp", "ego_lane", "target_lane", "n_points", "car_lane"):
        bad = np.argwhere(getattr(got, name) != getattr(want, name))
        assert len(bad) == 0, (name, bad[:5].tolist())
    assert np.array_equal(np.isnan(got.next_x), np.isnan(want.next_x))
    try:
        pp.set_kernel_variant(1)
        fused = gpu_plan(pp, torch_cuda, gmap, fb)
    finally:
        pp.set_kernel_variant(0)
    for name in ("next_x", "next_y", "n_points", "flags", "target_lane"):
        assert np.array_equal(getattr(got, name), getattr(fused, name), equal_nan=True), name


def test_plan_stats_batch_on_rare_and_poisoned_frames(pp, torch_cuda, gmap):
    """The fused statistics (checksum accumulated inside the planning kernels) against the
    stand-alone pass when many frames take the side paths — fallback generator, frames the
    emission kernel hands back, standstills with one NaN point, poisoned inputs — and more than
    one chunk is in flight."""
    n = 280000
    fb = pp.synth_frames(gmap, n, 6, seed=131, rare_permille=400)
    vals = [np.nan, np.inf, -np.inf, 1e200, 1e-310]
    fields = ["car_x", "car_y", "car_vx", "car_vy", "ego_yaw_deg", "ego_speed_mph", "prev_x", "prev_y"]
    for f in range(0, n, 97):
        a = getattr(fb, fields[(f // 5) % 8])
        if a.ndim == 2:
            a[f, (f // 40) % a.shape[1]] = vals[f % 5]
        else:
            a[f] = vals[f % 5]
    yaw = np.deg2rad(fb.ego_yaw_deg[1000:3000])
    fb.ego_speed_mph[1000:3000] = 0.0  # standstill behind a stopped car (one NaN point)
    fb.prev_n[1000:3000] = 0
    fb.n_cars[1000:3000] = 1
    with np.errstate(invalid="ignore"):  # a few of these yaws were poisoned above
        fb.car_x[1000:3000, 0] = fb.ego_x[1000:3000] + 14.7 * np.cos(yaw)
        fb.car_y[1000:3000, 0] = fb.ego_y[1000:3000] + 14.7 * np.sin(yaw)
    fb.car_vx[1000:3000, 0] = 0.0
    fb.car_vy[1000:3000, 0] = 0.0
    df = pp.DeviceFrames(fb)
    a = pp.DevicePlans(n, 6, diag=False, cars=False)
    b = pp.DevicePlans(n, 6, diag=False, cars=False)
    pp.plan_batch(gmap, df, a)
    want = pp.stats_batch(a).cpu().numpy()
    got = pp.plan_stats_batch(gmap, df, b).cpu().numpy()
    assert np.array_equal(got, want), (got, want)
    assert got[0] == n and got[1] < 50 * n  # some short paths are in there


@pytest.mark.parametrize("cars,rare", [(12, 300), (64, 100), (0, 50)])
def test_device_generator_is_the_host_generator(pp, torch_cuda, gmap, cars, rare):
    """pp_synth_frames_dev (one thread per frame, BASELINE config 5 generates its frames in HBM)
    writes the same bits as the host loop, rare frames included."""
    n, mc = 20011, max(cars, 1)
    host = pp.synth_frames(gmap, n, cars, seed=0xD00D, first_frame=123456789, rare_permille=rare,
                           max_cars=mc)
    dev = pp.synth_frames_dev(gmap, n, cars, seed=0xD00D, first_frame=123456789, rare_permille=rare,
                              max_cars=mc)
    torch_cuda.cuda.synchronize()
    got = dev.to_host()
    for k, v in host.arrays().items():
        assert np.array_equal(getattr(got, k), v, equal_nan=True), k


def numpy_fstats(pp, plans):
    """Mirror of pp_fstats_batch (definition in include/pp.h)."""
    x, y, npts = plans.next_x, plans.next_y, plans.n_points
    out = np.array([np.inf] * pp.FSTAT_NMIN + [-np.inf] * (pp.FSTATS_LEN - pp.FSTAT_NMIN))

    def upd(i, vals):
        vals = vals[np.isfinite(vals)]
        if len(vals):
            out[i] = min(out[i], vals.min()) if i < pp.FSTAT_NMIN else max(out[i], vals.max())
    upd(0, plans.ego_speed), upd(3, plans.ego_speed)
    upd(1, plans.target_speed), upd(4, plans.target_speed)
    with np.errstate(invalid="ignore", over="ignore"):
        vx, vy = (x[:, 1:] - x[:, :-1]) * 50, (y[:, 1:] - y[:, :-1]) * 50
        sp = np.sqrt(vx * vx + vy * vy)
        k = np.arange(sp.shape[1])[None, :]
        live = sp[k < (npts[:, None] - 1)]
        upd(2, live), upd(5, live)
        ax, ay = (vx[:, 1:] - vx[:, :-1]) * 50, (vy[:, 1:] - vy[:, :-1]) * 50
        acc = np.sqrt(ax * ax + ay * ay)
        k = np.arange(acc.shape[1])[None, :]
        upd(6, acc[k < (npts[:, None] - 2)])
    return out


def test_fstats_and_the_final_reduction(pp, torch_cuda, gmap):
    """f64 min / max statistics against a numpy mirror (bitwise: sqrt and + - * only), and
    pp_stats_reduce over a one-rank communicator made through the C ABI (the identity; the
    N-rank case runs in bench.py --gpus N, which asserts it against a single-rank pass)."""
    n = 50000
    fb = pp.synth_frames(gmap, n, 12, seed=88, rare_permille=200)
    df = pp.DeviceFrames(fb)
    dp = pp.DevicePlans(n, 12, diag=True, cars=False)
    st = pp.plan_stats_batch(gmap, df, dp)
    fs = pp.fstats_batch(dp)
    torch_cuda.cuda.synchronize()
    want = numpy_fstats(pp, dp.to_host())
    got = fs.cpu().numpy()
    assert np.array_equal(got, want), (got, want)
    assert got[pp.FSTAT_NMIN + 3] > 0 and got[0] >= 0  # some acceleration, speeds are norms
    before_i, before_f = st.clone(), fs.clone()
    comm = pp.Comm(0, 1)
    comm.stats_reduce(st, fs)
    torch_cuda.cuda.synchronize()
    assert torch_cuda.equal(st, before_i) and torch_cuda.equal(fs, before_f)
    comm.close()
