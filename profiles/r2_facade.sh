#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "facade or replay or split or multi_tool" > gpurun_out/facade_tests.log 2>&1; tail -3 gpurun_out/facade_tests.log
for i in 1 2; do
timeout 300 python bench.py --workload rollouts --rollouts 1048576 --ticks 60 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('rollouts 1M x 60:', d['value']/1e6, 'M ego-frames/s', d['config'])" | tee -a gpurun_out/r2_rollouts_1m.log
done
timeout 300 python bench.py --workload rollouts --rollouts 262144 --ticks 250 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('rollouts 256k x 250:', d['value']/1e6, 'M ego-frames/s', d['config'])" | tee -a gpurun_out/r2_rollouts_1m.log
