#!/bin/bash
# closed-loop rollouts: direct issue against graph replay, alternating, several runs each
mkdir -p gpurun_out
{
for rep in 1 2 3 4 5; do for mode in direct graph; do
  if [ $mode = graph ]; then export PP_ROLLOUT_GRAPH=1; else unset PP_ROLLOUT_GRAPH; fi
  echo "$mode: $(timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick" % (d["value"]/1e6, d["config"]["ms_per_tick"]))')"
done; done
} 2>&1 | tee gpurun_out/r2_rollouts_modes.log
