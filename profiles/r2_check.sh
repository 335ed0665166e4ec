#!/bin/bash
# Round-2 per-change check on the GPU box: the whole GPU suite, then per-phase timings of the
# default pipeline (2) and of the tiled-cars variant (3) at 12 and 64 cars per frame.
#   gpurun -- 'bash profiles/r2_check.sh TAG'
tag=${1:-x}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -6 gpurun_out/pytest_$tag.log
{
for v in 2 3; do
  echo "== variant $v, 12 cars"; PP_PIPES=1 python profiles/probe_overhead.py $v 1048576 2>&1 | tail -2; python profiles/probe_overhead.py $v 1048576 2>&1 | tail -1
  echo "== variant $v, 64 cars"; PP_PIPES=1 python profiles/probe_overhead.py $v 524288 64 2>&1 | tail -2; python profiles/probe_overhead.py $v 524288 64 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/probe_$tag.log
