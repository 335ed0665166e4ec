#!/bin/bash
mkdir -p gpurun_out
timeout 200 python bench.py --workload dense64 --steps 3 --e2e-steps 3 --no-cpu > gpurun_out/bench_r2j_dense64_n1.json 2> gpurun_out/bench_r2j_dense64_n1.err; tail -2 gpurun_out/bench_r2j_dense64_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2j_dense64_n1.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["whole_rows"]["value"], d["roofline"]["frac"])
PY
