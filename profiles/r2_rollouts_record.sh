#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
timeout 100 python bench.py --workload rollouts --no-cpu > gpurun_out/bench_r2j_rollouts_$rep.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2j_rollouts_$rep.json').read().strip().splitlines()[-1]); print('rollouts: %.1f M ego-frames/s, %.3f ms/tick, launches %d' % (d['value']/1e6, d['config']['ms_per_tick'], d['gpu_launches']))"
done
