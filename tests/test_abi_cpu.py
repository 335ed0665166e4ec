"""CPU-side checks of the product library: it loads, exports every symbol that
include/pp.h declares, Map::Init and the config defaults agree with the
reference, and — without a GPU — every planning call FAILS LOUDLY instead of
falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pp):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(pp.lib, n), f"libpp_b200.so does not export {n}"
    assert sorted(pp.EXPORTS) == names
    assert pp.lib.pp_version() == 102


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under the package may mention it."""
    pkg = os.path.join(ROOT, "carnd-path-planning-project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in text and "pporacle" not in text and "ppref" not in text, f
    hdr = open(os.path.join(ROOT, "include", "pp.hpp")).read() if os.path.exists(
        os.path.join(ROOT, "include", "pp.hpp")) else ""
    assert "oracle" not in hdr


def test_config_default_is_the_reference_literals(pp):
    cfg = pp.Config()
    assert pp.lib.pp_config_default(C.byref(cfg)) == 0
    want = pp.default_config()
    for name, _ in pp.Config._fields_:
        assert getattr(cfg, name) == getattr(want, name), name
    assert (cfg.relaxed_acc, cfg.min_relaxed_acc_while_braking, cfg.maximum_acc) == (5, 4, 8)
    assert (cfg.max_speed, cfg.car_length, cfg.safety_distance) == (22.2, 4.5, 2)
    assert (cfg.keep_distance, cfg.keep_distance_leeway, cfg.test_fast_lane_change) == (10, 0.5, 0)


def test_map_table_bit_identical_to_oracle(pp, pmap, oracle):
    assert pmap.n == 181
    assert np.array_equal(pmap.table(), oracle.map_table())


def test_map_from_points_and_errors(pp, oracle):
    csv = np.loadtxt(pp.MAP_CSV)
    m = pp.Map(points=(csv[:, 0], csv[:, 1]))
    assert np.array_equal(m.table(), oracle.map_table())
    h = C.c_void_p()
    assert pp.lib.pp_map_create_from_csv(b"/nonexistent/map.csv", C.byref(h)) == -3  # PP_E_IO
    assert pp.lib.pp_map_create(None, None, C.c_int(5), C.byref(h)) == -1  # PP_E_ARG
    one = np.zeros(1)
    assert pp.lib.pp_map_create(C.c_void_p(one.ctypes.data), C.c_void_p(one.ctypes.data),
                                C.c_int(1), C.byref(h)) == -5  # PP_E_RANGE
    assert pp.lib.pp_strerror(-2).decode().startswith("CUDA error")


def test_argument_validation(pp, pmap):
    cfg = pp.default_config()
    fb = pp.FrameBatch(4, 12)
    pb = pp.PlanBatch(4, 12)
    fs, ps = fb.struct(), pb.struct()
    lib = pp.lib
    assert lib.pp_plan_batch(None, C.byref(cfg), C.byref(fs), C.byref(ps), C.c_int64(4), None) == -1
    assert lib.pp_plan_batch(pmap.handle, C.byref(cfg), C.byref(fs), C.byref(ps), C.c_int64(-1),
                             None) == -1
    fs.max_cars = 65
    rc = lib.pp_plan_batch_host(pmap.handle, C.byref(cfg), C.byref(fs), C.byref(ps), C.c_int64(4))
    assert rc in (-5, -2)  # PP_E_RANGE (or PP_E_CUDA first when there is no device)
    assert lib.pp_set_kernel_variant(7) == -1
    assert lib.pp_spline_batch(None, None, 5, None, 1, None, C.c_int64(1), None) == -1
    # split rows: the four row arrays are required, whole rows may not be asked for as well
    fs.max_cars = 12
    z = np.zeros((4, 50))
    rows = pp.abi.SplitRows(z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data)
    args = (pmap.handle, C.byref(cfg), C.byref(fs), C.byref(ps))
    assert lib.pp_plan_batch_host_split(*args, None, C.c_int64(4)) == -1
    assert lib.pp_plan_batch_host_split(*args, C.byref(rows), C.c_int64(4)) == -1  # next_x set
    ps.next_x = ps.next_y = None
    rows.tail_y = None
    assert lib.pp_plan_batch_host_split(*args, C.byref(rows), C.c_int64(4)) == -1
    assert lib.pp_host_alloc(None, C.c_size_t(16)) == -1
    assert lib.pp_host_free(None) == 0


def test_no_cpu_fallback_without_gpu(pp, pmap):
    """On a box without CUDA the product must refuse to plan (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present: covered by the gpu tests")
    assert pp.device_count() < 0 or pp.device_count() == 0
    fb = pp.synth_frames(pmap, 8, 12)
    with pytest.raises(pp.PPError, match="CUDA"):
        pp.plan_batch_host(pmap, fb)
    sp = pp.PlanBatch(8, 12, diag=False, cars=False)
    sp.next_x = sp.next_y = None
    with pytest.raises(pp.PPError, match="CUDA"):
        pp.plan_batch_host_split(pmap, fb, sp, fb.prev_x, fb.prev_y, np.zeros((8, 40)),
                                 np.zeros((8, 40)))
    with pytest.raises(pp.PPError, match="CUDA"):
        pp.pinned_empty(16)
    assert pp.launch_count() == 0


def test_synthetic_generator_is_counter_based(pp, pmap):
    """Any sub-range equals the same slice of the whole (what lets ranks
    generate their own shards), and seeds / car counts change the data."""
    whole = pp.synth_frames(pmap, 300, 12, seed=5)
    part = pp.synth_frames(pmap, 100, 12, seed=5, first_frame=150)
    for k, v in whole.arrays().items():
        assert np.array_equal(v[150:250], getattr(part, k)), k
    other = pp.synth_frames(pmap, 300, 12, seed=6)
    assert not np.array_equal(whole.prev_x, other.prev_x)
    assert whole.bytes_per_frame() == 204 + 36 * 12  # SURVEY §8d algorithmic input bytes
    dense = pp.synth_frames(pmap, 10, 64)
    assert dense.bytes_per_frame() == 204 + 36 * 64
    assert (dense.n_cars == 64).all() and dense.car_id[0, 63] in (0, 63)
    pb = pp.PlanBatch(1, 12, diag=True, cars=False)
    assert pb.bytes_per_frame() == 884 + 8  # SURVEY §8d outputs + the two followed-car ids


def test_synthetic_frames_are_plannable(pp, pmap, oracle):
    """The generator's promise: strictly increasing local x for ordinary frames,
    i.e. the spline path (not the fallback) is taken, and rare frames do reach
    the rare branches."""
    fb = pp.synth_frames(pmap, 4000, 12, seed=9, rare_permille=0)
    p = oracle.plan(fb)
    assert (p.flags & pp.FLAG["FALLBACK"]).astype(bool).mean() < 0.01
    assert (p.n_points == 50).all()
    fb = pp.synth_frames(pmap, 4000, 12, seed=9, rare_permille=1000)
    p = oracle.plan(fb)
    for name in ("EGO_MATCH_FAIL", "CAR_DROPPED", "COLLISION", "BRAKE", "MAXBRAKE", "ADJUST", "KEEP",
                 "SPLINE_INPUT_ERR", "FALLBACK", "ACC_OVERRIDE", "CURV_ADJUST", "VETO",
                 "COLD_START"):
        assert (p.flags & pp.FLAG[name]).any(), name
