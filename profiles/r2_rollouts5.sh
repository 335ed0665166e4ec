#!/bin/bash
# does the NVML clock sampler (a thread polling every 2 ms) disturb the launch-bound rollout job?
mkdir -p gpurun_out
{
for rep in 1 2 3 4; do for ms in 2 50; do
  echo "sampler ${ms} ms: $(PP_BENCH_CLOCK_MS=$ms timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, clocks %s" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["clocks"]))')"
done; done
} 2>&1 | tee gpurun_out/r2_rollouts_sampler.log
