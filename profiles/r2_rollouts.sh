#!/bin/bash
# closed-loop rollouts (configs[2]): stream groups x direct issue / graph replay
mkdir -p gpurun_out
{
for g in 4 8 2; do for mode in direct graph; do
  if [ $mode = graph ]; then export PP_ROLLOUT_GRAPH=1; else unset PP_ROLLOUT_GRAPH; fi
  for rep in 1 2; do
    echo "groups $g $mode: $(PP_ROLLOUT_GROUPS=$g timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
  done
done; done
} 2>&1 | tee gpurun_out/r2_rollouts.log
