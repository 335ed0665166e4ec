"""include/pp_wire.hpp — host-side logic only (no GPU): telemetry parsing, replies, and the
cross-frame state of a simulator session (the reference's persistent car map and target_lane,
src/main.cpp:1194-1195,1217-1252,1325-1340,1461-1471).  tests/cpp/test_wire.cpp holds the
cases; the end-to-end comparison with the untouched onMessage lambda is tests/test_replay_tool.py."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "carnd-path-planning-project_b200")
SRC = os.path.join(ROOT, "tests", "cpp", "test_wire.cpp")


def test_wire_codec_and_session_state(tmp_path, pp):
    exe = str(tmp_path / "test_wire")
    cmd = ["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", PKG, "-lpp_b200", "-Wl,-rpath," + PKG, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    res = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0 and "wire ok" in res.stdout, (res.stdout, res.stderr)
