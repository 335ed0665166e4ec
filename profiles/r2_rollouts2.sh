#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollouts.py -x -q -m gpu > gpurun_out/rollout_tests.log 2>&1; tail -3 gpurun_out/rollout_tests.log
for i in 1 2 3 4; do
timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('rollouts 65536 x 1000:', round(d['value']/1e6,1), 'M ego-frames/s', d['config']['ms_per_tick'])" | tee -a gpurun_out/r2_rollouts2.log
done
timeout 300 python bench.py --workload rollouts --rollouts 1048576 --ticks 60 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('rollouts 1M x 60:', round(d['value']/1e6,1), 'M ego-frames/s', d['config']['ms_per_tick'])" | tee -a gpurun_out/r2_rollouts2.log
