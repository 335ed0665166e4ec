#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollouts.py -x -q -m gpu > gpurun_out/rollout_tests.log 2>&1; tail -3 gpurun_out/rollout_tests.log
{
for rep in 1 2 3; do
  echo "${1:-x}: $(timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
done
echo "${1:-x} 1 group: $(PP_ROLLOUT_GROUPS=1 timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
echo "${1:-x} 1M: $(timeout 300 python bench.py --workload rollouts --rollouts 1048576 --ticks 60 --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
} 2>&1 | tee -a gpurun_out/r2_rollouts_sim2.log
