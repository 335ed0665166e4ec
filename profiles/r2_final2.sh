#!/bin/bash
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for rep in 1 2 3; do
timeout 300 python bench.py --workload rollouts --no-cpu > gpurun_out/bench_${tag}_rollouts_$rep.json 2>/dev/null
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${tag}_rollouts_$rep.json').read().strip().splitlines()[-1]); print('rollouts: %.1f M ego-frames/s, %.3f ms/tick, launches %d' % (d['value']/1e6, d['config']['ms_per_tick'], d['gpu_launches']))"
done
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 300 gpurun_out/bench_${tag}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["mean_value"], d["e2e"]["whole_rows"]["value"], d["clocks"], d["gpu_launches"])
PY
