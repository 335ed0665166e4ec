#!/bin/bash
# stream groups again, with the simulator kernels spread out, the checksum taken by the planning
# kernels and the clock sampler at 20 ms: direct issue, four runs each
mkdir -p gpurun_out
{
for rep in 1 2 3 4; do for groups in 1 2 4 8; do
  echo "groups $groups: $(PP_ROLLOUT_GROUPS=$groups timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d, samples %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"], d["clocks"]["samples"]))')"
done; done
} 2>&1 | tee gpurun_out/r2_rollouts_groups2.log
