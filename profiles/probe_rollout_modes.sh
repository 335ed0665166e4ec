#!/bin/bash
# Run-to-run spread of the closed-loop rollout job, direct issue against graph replay.
for i in 1 2 3; do
python bench.py --workload rollouts --no-cpu "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('direct', round(d['value']/1e6,1), round(d['ms_per_step'],1), d['clocks'])"
PP_ROLLOUT_GRAPH=1 python bench.py --workload rollouts --no-cpu "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graph ', round(d['value']/1e6,1), round(d['ms_per_step'],1), d['clocks'])"
done
uptime
