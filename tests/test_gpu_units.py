"""Unit-level GPU parity: one batch kernel per reference function (the same
__device__ code the fused kernel runs), each against the committed golden
vectors produced by the reference's own functions and against the oracle.
Everything here except the transcendental-based starter helpers is bit-exact."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


@pytest.fixture(scope="module")
def gmap(pp, T):
    return pp.Map()


def dev(T, a, dtype=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    return T.from_numpy(a).cuda()


def p(t):
    return C.c_void_p(t.data_ptr())


def out_f64(T, *shape):
    return T.zeros(shape, dtype=T.float64, device="cuda")


def out_i32(T, *shape):
    return T.zeros(shape, dtype=T.int32, device="cuda")


def test_distancesq_pt_seg(pp, T, golden_units, oracle):
    g = golden_units
    ins = [dev(T, a) for a in g["seg_in"]]
    n = ins[0].numel()
    outs = [out_f64(T, n) for _ in range(4)]
    rc = pp.lib.pp_distancesq_pt_seg_batch(*[p(t) for t in ins], *[p(t) for t in outs],
                                           C.c_int64(n), None)
    assert rc == 0
    got = np.stack([o.cpu().numpy() for o in outs])
    assert np.array_equal(got, g["seg_out"])
    # reference quirks are present in the vectors
    assert (g["seg_out"][1][8:16] < 0).all()            # -1 <= rnom < 0 left un-clamped
    assert (g["seg_out"][0][:8] == 0).all() and (g["seg_out"][2][:8] == 1).all()  # degenerate


def test_init_reference_waypoint(pp, T, gmap, golden_units):
    g = golden_units
    x, y = (dev(T, a) for a in g["ref_in"])
    n = x.numel()
    wp, ratio = out_i32(T, n), out_f64(T, n, 3)
    assert pp.lib.pp_init_reference_waypoint_batch(gmap.handle, p(x), p(y), p(wp), p(ratio),
                                                   C.c_int64(n), None) == 0
    assert np.array_equal(wp.cpu().numpy(), g["ref_wp"])
    assert np.array_equal(ratio.cpu().numpy(), g["ref_ratio"])


def test_lane_matching_and_project_speed(pp, T, gmap, golden_units):
    g = golden_units
    ins = [dev(T, a) for a in g["lm_in"]]
    n = ins[0].numel()
    ok, lane, nwp = (out_i32(T, n) for _ in range(3))
    s, d, vs, vd = (out_f64(T, n) for _ in range(4))
    assert pp.lib.pp_lane_matching_batch(gmap.handle, *[p(t) for t in ins], p(ok), p(lane), p(nwp),
                                         p(s), p(d), p(vs), p(vd), C.c_int64(n), None) == 0
    got = dict(ok=ok, lane=lane, next_wp=nwp, s=s, d=d, vs=vs, vd=vd)
    for k, v in got.items():
        assert np.array_equal(v.cpu().numpy(), g["lm_" + k]), k


def test_get_lane_pos(pp, T, gmap, golden_units):
    g = golden_units
    rx, ry = (dev(T, a) for a in g["ref_in"])
    s, lane = dev(T, g["lp_s"]), dev(T, g["lp_lane"], np.int32)
    n = s.numel()
    ox, oy, od, owp = out_f64(T, n), out_f64(T, n), out_f64(T, n), out_i32(T, n)
    assert pp.lib.pp_get_lane_pos_batch(gmap.handle, p(rx), p(ry), p(s), p(lane), p(ox), p(oy),
                                        p(owp), p(od), C.c_int64(n), None) == 0
    assert np.array_equal(ox.cpu().numpy(), g["lp_x"]) and np.array_equal(oy.cpu().numpy(), g["lp_y"])
    assert np.array_equal(owp.cpu().numpy(), g["lp_wp"])
    assert np.array_equal(od.cpu().numpy(), g["lp_dist"])


@pytest.mark.parametrize("nk", [3, 6, 13, 15])
def test_spline(pp, T, golden_units, nk):
    g = golden_units
    kx, ky, q = dev(T, g[f"sp{nk}_kx"]), dev(T, g[f"sp{nk}_ky"]), dev(T, g[f"sp{nk}_q"])
    ns, nq = q.shape
    out = out_f64(T, ns, nq)
    assert pp.lib.pp_spline_batch(p(kx), p(ky), nk, p(q), nq, p(out), C.c_int64(ns), None) == 0
    assert np.array_equal(out.cpu().numpy(), g[f"sp{nk}_out"])  # bit-exact: only + - * /


def test_spline_rejects_bad_sizes(pp, T):
    z = out_f64(T, 4)
    assert pp.lib.pp_spline_batch(p(z), p(z), 2, p(z), 1, p(z), C.c_int64(1), None) == -5
    assert pp.lib.pp_spline_batch(p(z), p(z), 16, p(z), 1, p(z), C.c_int64(1), None) == -5


def test_starter_frenet_helpers(pp, T, golden_units):
    """ClosestWaypoint / NextWaypoint / getFrenet / getXY (src/helpers.h:43-155)."""
    g = golden_units
    csv = np.loadtxt(pp.MAP_CSV)
    mx, my, ms = dev(T, csv[:, 0]), dev(T, csv[:, 1]), dev(T, csv[:, 2])
    x, y, th = (dev(T, a) for a in g["hw_xyth"])
    n, nwp = x.numel(), mx.numel()
    o = out_i32(T, n)
    assert pp.lib.pp_closest_waypoint_batch(p(x), p(y), p(mx), p(my), nwp, p(o), C.c_int64(n), None) == 0
    assert np.array_equal(o.cpu().numpy(), g["hw_closest"])
    assert pp.lib.pp_next_waypoint_batch(p(x), p(y), p(th), p(mx), p(my), nwp, p(o), C.c_int64(n),
                                         None) == 0
    # NextWaypoint compares an atan2 result with pi/2: allow the (measure-zero) ulp flips
    assert (o.cpu().numpy() != g["hw_next"]).mean() < 0.005
    fs, fd = out_f64(T, n), out_f64(T, n)
    assert pp.lib.pp_get_frenet_batch(p(x), p(y), p(th), p(mx), p(my), nwp, p(fs), p(fd),
                                      C.c_int64(n), None) == 0
    same = o.cpu().numpy() == g["hw_next"]
    got = np.stack([fs.cpu().numpy(), fd.cpu().numpy()])
    assert np.array_equal(got[:, same], g["hw_frenet"][:, same])  # only + - * / sqrt after the index
    s, d = (dev(T, a) for a in g["hw_sd"])
    ox, oy = out_f64(T, n), out_f64(T, n)
    assert pp.lib.pp_get_xy_batch(p(s), p(d), p(ms), p(mx), p(my), nwp, p(ox), p(oy), C.c_int64(n),
                                  None) == 0
    got = np.stack([ox.cpu().numpy(), oy.cpu().numpy()])
    assert np.allclose(got, g["hw_xy"], rtol=1e-12, atol=1e-9)  # atan2/cos/sin inside


def test_lane_change_planner(pp, T, golden_units):
    """LaneChangePlanner::calculate_target_lane (src/main.cpp:364-485), bit-exact,
    incl. ties between two cars (lowest id wins) and all three target lanes."""
    g = golden_units
    cs, cvs = g["lc_cars"]
    es, evs, dt0 = g["lc_ego"]
    n, nc = g["lc_id"].shape
    cfg = pp.default_config()
    out = out_i32(T, n)
    args = [dev(T, g["lc_id"], np.int32), dev(T, cs), dev(T, cvs), dev(T, g["lc_lane"], np.int32)]
    sc = [dev(T, g["lc_el"], np.int32), dev(T, g["lc_tl"], np.int32), dev(T, es), dev(T, evs),
          dev(T, dt0)]
    assert pp.lib.pp_lane_change_batch(C.byref(cfg), *[p(t) for t in args], nc, *[p(t) for t in sc],
                                       p(out), C.c_int64(n), None) == 0
    assert np.array_equal(out.cpu().numpy(), g["lc_out"])
    # car order must not matter
    perm = np.random.default_rng(1).permutation(nc)
    args = [dev(T, g["lc_id"][:, perm], np.int32), dev(T, cs[:, perm]), dev(T, cvs[:, perm]),
            dev(T, g["lc_lane"][:, perm], np.int32)]
    assert pp.lib.pp_lane_change_batch(C.byref(cfg), *[p(t) for t in args], nc, *[p(t) for t in sc],
                                       p(out), C.c_int64(n), None) == 0
    assert np.array_equal(out.cpu().numpy(), g["lc_out"])


def test_limit_speed_and_speed_controller(pp, T, golden_units):
    """LimitSpeed::calculate + SpeedController::add_limit_breakpoint, bit-exact."""
    g = golden_units
    ins = [dev(T, a) for a in g["ls_in"]]
    inl = dev(T, g["ls_inlane"], np.int32)
    n = inl.numel()
    outs = [out_f64(T, n) for _ in range(4)]
    fl = out_i32(T, n)
    cfg = pp.default_config()
    assert pp.lib.pp_limit_speed_batch(C.byref(cfg), *[p(t) for t in ins], p(inl),
                                       *[p(t) for t in outs], p(fl), C.c_int64(n), None) == 0
    for o, k in zip(outs, ("ls_speed", "ls_time", "sc_speed", "sc_time")):
        assert np.array_equal(o.cpu().numpy(), g["ls_" + k]), k
    assert np.array_equal(fl.cpu().numpy().view(np.uint32), g["ls_flags"])
    for name in ("COLLISION", "BRAKE", "MAXBRAKE", "ADJUST", "KEEP"):
        assert (g["ls_flags"] & pp.FLAG[name]).any(), name


def test_trajectory_builder(pp, T, gmap, golden_units):
    """TrajectoryBuilder::build on explicit inputs vs the reference's own class:
    point counts identical, points within 1e-9 rel / 1e-6 m."""
    g = golden_units
    ex, ey, yaw, ed, evd, start, target, time = g["tr_ego"]
    n = len(ex)
    cfg = pp.default_config()
    ox, oy = out_f64(T, n, 50), out_f64(T, n, 50)
    on, fl = out_i32(T, n), out_i32(T, n)
    a = [dev(T, g["tr_prev_n"], np.int32), dev(T, g["tr_prev_x"]), dev(T, g["tr_prev_y"]), dev(T, ex),
         dev(T, ey), dev(T, yaw), dev(T, g["tr_tl"], np.int32), dev(T, ed), dev(T, evd), dev(T, start),
         dev(T, target), dev(T, time)]
    assert pp.lib.pp_trajectory_build_batch(gmap.handle, C.byref(cfg), *[p(t) for t in a], p(ox),
                                            p(oy), p(on), p(fl), C.c_int64(n), None) == 0
    assert np.array_equal(on.cpu().numpy(), g["tr_n"])
    for got, want in ((ox.cpu().numpy(), g["tr_x"]), (oy.cpu().numpy(), g["tr_y"])):
        assert np.array_equal(got != got, want != want)
        err = np.nanmax(np.abs(got - want))
        assert err <= 1e-6 and np.nanmax(np.abs(got - want) / np.abs(want)) <= 1e-9, err


def test_exact_arithmetic_helpers_selftest(pp, T):
    """The Markstein quotients (cached reciprocal, x/50) and the Sterbenz angle
    wrap must be bit-identical to the generic / and fmod they replace — 2^27
    random trials incl. extreme magnitudes — and the forward-step atan (slopes up
    to +-3, i.e. all three rotation classes and the library fall-through) must stay
    within 6 ulp of the library atan2 (< 7e-16 rad; trajectory tolerance 1e-9)."""
    counts = T.zeros(8, dtype=T.int64, device="cuda")
    assert pp.lib.pp_selftest_math(C.c_int64(1 << 27), C.c_uint64(12345), p(counts), None) == 0
    c = counts.cpu().numpy()
    assert c[0] == 0 and c[1] == 0 and c[2] == 0, c
    assert 0 <= c[3] <= 6, c
    # the emission kernel's unguarded versions (reciprocal without range checks, one-correction
    # x/50, atan and angle wrap without fall-backs) equal the guarded helpers bit for bit
    assert (c[4:8] == 0).all(), c
