#!/bin/bash
# closed-loop rollouts after the simulator kernels were spread over 8 threads per rollout:
# stream groups x (direct issue | graph replay), three runs each
mkdir -p gpurun_out
{
for groups in 8 4 2 1; do for mode in direct graph; do for rep in 1 2 3; do
  if [ $mode = graph ]; then export PP_ROLLOUT_GRAPH=1; else unset PP_ROLLOUT_GRAPH; fi
  echo "groups $groups $mode: $(PP_ROLLOUT_GROUPS=$groups timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
done; done; done
} 2>&1 | tee gpurun_out/r2_rollouts_matrix.log
