#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU suite, smoke(), both bench arms
tag=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; tail -c 400 gpurun_out/bench_${tag}_reference.json
timeout 900 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 300 gpurun_out/bench_${tag}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["mean_value"], d["e2e"]["whole_rows"]["value"], d["clocks"], d["gpu_launches"])
PY
