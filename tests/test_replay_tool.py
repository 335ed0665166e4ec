"""tools/pp_replay + include/pp_wire.hpp: recorded simulator messages ->
frames -> GPU planner -> the replies the reference would send, and the trajectory.log header.

GPU: a session of 42["telemetry",{...}] messages is replayed by the tool and by the
reference's UNTOUCHED onMessage lambda (oracle/_ref lambda harness: its own JSON parsing,
its own persistent target_lane); the replies must carry the same points (both sides print 15
significant digits).  CPU: the tool builds as plain C++11 and fails loudly without a GPU.
"""
import json
import os
import subprocess

import numpy as np
import pytest

import checkers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "carnd-path-planning-project_b200")
CSV = os.path.join(ROOT, "data", "highway_map.csv")


@pytest.fixture(scope="module")
def replay_exe(tmp_path_factory, pp):
    exe = str(tmp_path_factory.mktemp("replay") / "pp_replay")
    cmd = ["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tools", "pp_replay.cpp"), "-L", PKG, "-lpp_b200",
           "-Wl,-rpath," + PKG, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def telemetry_message(fb, f):
    """The message the simulator would send for frame f (src/main.cpp:1233-1252,1297)."""
    pn = int(fb.prev_n[f])
    px = [float(fb.prev_x[f, min(i, 9)]) for i in range(pn)]   # only the first 10 matter
    py = [float(fb.prev_y[f, min(i, 9)]) for i in range(pn)]
    sf = [[int(fb.car_id[f, j]), float(fb.car_x[f, j]), float(fb.car_y[f, j]),
           float(fb.car_vx[f, j]), float(fb.car_vy[f, j]), 0, 0] for j in range(int(fb.n_cars[f]))]
    d = {"x": float(fb.ego_x[f]), "y": float(fb.ego_y[f]), "s": 0, "d": 0,
         "yaw": float(fb.ego_yaw_deg[f]), "speed": float(fb.ego_speed_mph[f]),
         "previous_path_x": px, "previous_path_y": py, "end_path_s": 0, "end_path_d": 0,
         "sensor_fusion": sf}
    return '42["telemetry",' + json.dumps(d) + "]"


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_replay_tool_builds_and_fails_loudly_without_gpu(replay_exe, pp, tmp_path):
    if _has_gpu():
        pytest.skip("a GPU is present: covered by the gpu test")
    fb = pp.synth_frames(pp.Map(), 2, 12, seed=1, rare_permille=0)
    session = tmp_path / "s.txt"
    session.write_text("\n".join(telemetry_message(fb, f) for f in range(2)) + "\n")
    res = subprocess.run([replay_exe, "--map", CSV, str(session)], capture_output=True, text=True)
    assert res.returncode == 2 and "pp::Error" in res.stderr


@pytest.mark.gpu
def test_replay_matches_the_untouched_reference_lambda(replay_exe, pp, ref, tmp_path):
    n = 120
    fb = pp.synth_frames(pp.Map(), n, 12, seed=909, rare_permille=60)
    want_x, want_y, want_n = ref.lambda_sequence(fb)
    lines = [telemetry_message(fb, f) for f in range(n)]
    lines.insert(40, '42["telemetry",null]')           # manual driving (hasData -> "")
    session = tmp_path / "session.txt"
    session.write_text("\n".join(lines) + "\n")
    log = tmp_path / "trajectory.log"
    res = subprocess.run([replay_exe, "--map", CSV, "--log", str(log), str(session)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    replies = (tmp_path / "session.txt.out").read_text().splitlines()
    assert len(replies) == n + 1 and replies[40] == '42["manual",{}]'
    del replies[40]
    for f, r in enumerate(replies):
        assert r.startswith('42["control",')
        body = json.loads(r[2:])[1]
        gx, gy = np.array(body["next_x"], dtype=float), np.array(body["next_y"], dtype=float)
        assert len(gx) == len(gy) == want_n[f], f
        assert np.allclose(gx, want_x[f, :want_n[f]], rtol=1e-9, atol=1e-6, equal_nan=True), f
        assert np.allclose(gy, want_y[f, :want_n[f]], rtol=1e-9, atol=1e-6, equal_nan=True), f
    # trajectory.log: the header the reference writes (DrawLines.ipynb reads these arrays)
    head = log.read_text().splitlines()
    assert head[0].startswith("wpmap=[[784.6001,1135.5710]") and head[1].startswith("lane0=[[784.55")
    tbl = pp.Map().table()
    lane2 = json.loads(head[3].split("=", 1)[1])
    assert np.allclose(np.array(lane2), tbl[:, 6:8], atol=5.1e-5)
    assert sum(1 for ln in head if ln.startswith("result=[")) == n


@pytest.mark.gpu
def test_replay_keeps_the_reference_persistent_car_map(replay_exe, pp, ref, abi, tmp_path):
    """The reference's sensor_fusion_cars map outlives the frame (src/main.cpp:1194,1325-1340):
    a car missing from a message keeps planning with the values of its last sighting, a repeated
    id keeps its last row, a car that fails lane matching is erased.  A session in which ids
    drop out, come back and repeat must match the untouched lambda, which carries that map."""
    n = 160
    src = pp.synth_frames(pp.Map(), n, 12, seed=1234, rare_permille=120)
    rng = np.random.default_rng(12)
    fb = abi.FrameBatch(n, 14)
    for k in ("ego_x", "ego_y", "ego_yaw_deg", "ego_speed_mph", "prev_n", "prev_x", "prev_y",
              "target_lane_in"):
        getattr(fb, k)[:] = getattr(src, k)
    for f in range(n):
        keep = np.flatnonzero(rng.random(12) < (1.0 if f % 7 == 0 else 0.6))  # ids drop out / return
        rows = list(keep)
        if f % 5 == 3 and len(rows):  # a repeated id: the second row (another car's place) wins
            rows.append(rows[0])
        fb.n_cars[f] = len(rows)
        for j, r in enumerate(rows):
            pos = r if j < len(keep) else (r + 5) % 12
            fb.car_id[f, j] = src.car_id[f, r]
            fb.car_x[f, j], fb.car_y[f, j] = src.car_x[f, pos], src.car_y[f, pos]
            fb.car_vx[f, j], fb.car_vy[f, j] = src.car_vx[f, pos], src.car_vy[f, pos]
    want_x, want_y, want_n = ref.lambda_sequence(fb)
    session = tmp_path / "session.txt"
    session.write_text("\n".join(telemetry_message(fb, f) for f in range(n)) + "\n")
    res = subprocess.run([replay_exe, "--map", CSV, str(session)], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr
    replies = (tmp_path / "session.txt.out").read_text().splitlines()
    assert len(replies) == n
    for f, r in enumerate(replies):
        body = json.loads(r[2:])[1]
        gx, gy = np.array(body["next_x"], dtype=float), np.array(body["next_y"], dtype=float)
        assert len(gx) == len(gy) == want_n[f], f
        assert np.allclose(gx, want_x[f, :want_n[f]], rtol=1e-9, atol=1e-6, equal_nan=True), f
        assert np.allclose(gy, want_y[f, :want_n[f]], rtol=1e-9, atol=1e-6, equal_nan=True), f
    # ... and it does matter: without the held-over cars some plans differ
    stateless = [pp.plan_batch_host(pp.Map(), fb.slice(f, f + 1)) for f in range(1, n)]
    diff = sum(1 for f, p in enumerate(stateless, start=1)
               if not np.allclose(p.next_x[0, :want_n[f]], want_x[f, :want_n[f]], rtol=1e-9,
                                  atol=1e-6, equal_nan=True))
    assert diff > 0


@pytest.mark.gpu
def test_trajectory_log_control_points_match_the_reference_log(replay_exe, pp, ref, tmp_path):
    """trajectory.log: the builder's control points (control_points=, src/main.cpp:779-781) and the
    "first control dist" line against the text the reference's own classes write for the same
    frames (4 decimals, as DrawLines.ipynb reads them)."""
    import re
    n = 90
    fb = pp.synth_frames(pp.Map(), n, 12, seed=777, rare_permille=100)
    session = tmp_path / "session.txt"
    session.write_text("\n".join(telemetry_message(fb, f) for f in range(n)) + "\n")
    log = tmp_path / "trajectory.log"
    res = subprocess.run([replay_exe, "--map", CSV, "--log", str(log), str(session)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr
    ours = log.read_text().splitlines()
    # the replay carries target_lane from plan to plan (:1195): give the class harness the same
    tl_out = [int(m.group(1)) for ln in ours for m in [re.match(r"flags=0x[0-9a-f]+ target_lane=(\d)", ln)] if m]
    assert len(tl_out) == n
    fb.target_lane_in[0] = 1
    fb.target_lane_in[1:] = tl_out[:-1]
    ref_log = tmp_path / "ref.log"
    ref.plan_with_log(fb, str(ref_log))
    theirs = ref_log.read_text().splitlines()

    def arrays(lines, key):
        return [np.array(json.loads(re.sub(r"-?nan", "NaN", ln.split("=", 1)[1])), dtype=float).reshape(-1, 2)
                for ln in lines if ln.startswith(key + "=[")]
    a, b = arrays(ours, "control_points"), arrays(theirs, "control_points")
    assert len(a) == len(b) == n
    for f in range(n):
        assert a[f].shape == b[f].shape, (f, a[f].shape, b[f].shape)
        assert np.allclose(a[f], b[f], atol=1.01e-4, equal_nan=True), f
    da = [float(ln.split()[-1]) for ln in ours if ln.startswith("first control dist")]
    db = [float(ln.split()[-1]) for ln in theirs if ln.startswith("first control dist")]
    assert len(da) == len(db) == n and np.allclose(da, db, atol=0.0101)
    ra, rb = arrays(ours, "result"), arrays(theirs, "result")
    assert len(ra) == len(rb) == n
    for f in range(n):
        assert ra[f].shape == rb[f].shape and np.allclose(ra[f], rb[f], atol=1.01e-4, equal_nan=True), f
