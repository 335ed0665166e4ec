#!/bin/bash
mkdir -p gpurun_out
{
for rep in 1 2 3; do
echo "${1:-x} 1M: $(timeout 300 python bench.py --workload rollouts --rollouts 1048576 --ticks 60 --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d %s" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"], d["clocks"]))')"
done
echo "${1:-x} 256k: $(timeout 300 python bench.py --workload rollouts --rollouts 262144 --ticks 250 --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
} 2>&1 | tee -a gpurun_out/r2_rollouts_sim2.log
