#!/bin/bash
# Round-2 exploratory timings on the GPU box (per-phase ms with one chunk at a time, then the
# overlapped total) for a list of environment settings, one per argument ("A=1 B=2").
#   gpurun -- 'bash profiles/r2_probe.sh TAG "PP_FRONT_THREADS=512" "PP_FRONT_THREADS=640 PP_EMIT_TILE_OUT=1"'
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "variants or geometries or unaligned" > gpurun_out/pytest_new_$tag.log 2>&1; tail -3 gpurun_out/pytest_new_$tag.log
{
echo "== variant 2"; PP_PIPES=1 python profiles/probe_overhead.py 2 1048576 2>&1 | tail -2; python profiles/probe_overhead.py 2 1048576 2>&1 | tail -1
for cfg in "" "$@"; do
  echo "== variant 3 [$cfg]"
  env $cfg PP_PIPES=1 python profiles/probe_overhead.py 3 1048576 2>&1 | tail -2
  env $cfg python profiles/probe_overhead.py 3 1048576 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/probe_$tag.log
