"""Build recipe for libpp_b200.so (in-tree, sm_100a only).

    python carnd-path-planning-project_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  -fmad=false is part of the numerical
contract (no FMA contraction: bit-identical discrete results),
not a tuning knob.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpp_b200.so")
SOURCES = ["pp_api.cu", "pp_plan.cu", "pp_units.cu", "pp_rollout.cu", "pp_sweep.cu", "pp_map_host.cpp",
           "pp_synth.cu", "pp_comm.cu"]
HEADERS = ["pp_device.cuh", "pp_internal.h", os.path.join("..", "..", "include", "pp.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-shared",
    "-ldl",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("PP_EXTRA_NVCC_FLAGS", "").split()  # tuning experiments only
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
