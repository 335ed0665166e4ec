#!/bin/bash
# Run on the GPU box: parity tests, then the pipeline timing probe (per-phase ms).
#   gpurun -- 'bash profiles/gpu_check.sh TAG'
tag=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
python profiles/probe_overhead.py 0 1048576 2>&1 | tail -2 | tee gpurun_out/probe_$tag.log
