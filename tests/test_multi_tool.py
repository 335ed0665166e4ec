"""tools/pp_multi.cpp — the multi-GPU job as one C++ host program (one thread per device)
through the C ABI alone: contiguous shards generated in HBM, planned, and reduced once with
pp_stats_reduce (NCCL).  CPU: it builds as plain C++11 and fails loudly without a GPU.  GPU:
one device here (the N-device run and its --check are part of profiles/)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "carnd-path-planning-project_b200")
CSV = os.path.join(ROOT, "data", "highway_map.csv")


@pytest.fixture(scope="module")
def multi_exe(tmp_path_factory, pp):
    exe = str(tmp_path_factory.mktemp("multi") / "pp_multi")
    cmd = ["g++", "-std=c++11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tools", "pp_multi.cpp"), "-L", PKG, "-lpp_b200",
           "-Wl,-rpath," + PKG, "-pthread", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_multi_tool_builds_and_fails_loudly_without_gpu(multi_exe):
    if _has_gpu():
        pytest.skip("a GPU is present: covered by the gpu test")
    res = subprocess.run([multi_exe, CSV, "--frames", "1000"], capture_output=True, text=True,
                         timeout=120)
    assert res.returncode == 2, (res.returncode, res.stdout, res.stderr)
    assert "no CPU planning path" in res.stderr


@pytest.mark.gpu
def test_multi_tool_on_the_visible_devices(multi_exe, pp):
    import torch
    import numpy as np
    res = subprocess.run([multi_exe, CSV, "--frames", "300001", "--cars", "12", "--steps", "2",
                          "--chunk", "131072", "--check"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout, res.stderr)
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["stat_frames"] == 300001 and line["frames_per_s"] > 1e6
    if line["gpus"] > 1:
        assert "==" in res.stdout
    # the tool's totals are those of pp_plan_stats_batch over the same frames in this process
    torch.cuda.set_device(0)
    m = pp.Map()
    df = pp.synth_frames_dev(m, 300001, 12, seed=0x5EED, first_frame=0)
    dp = pp.DevicePlans(300001, 12, diag=True, cars=False)
    st = pp.plan_stats_batch(m, df, dp).cpu().numpy()
    fs = pp.fstats_batch(dp).cpu().numpy()
    assert int(st[1]) == line["stat_points"]
    assert np.float64(line["max_acc"]) == fs[pp.FSTAT_NMIN + 3]
