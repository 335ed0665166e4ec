#!/bin/bash
# split-rows host entry point: parity tests, then the bench line (e2e: split rows and whole rows)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "split_rows or host_entry" > gpurun_out/split_tests.log 2>&1
tail -3 gpurun_out/split_tests.log
log=gpurun_out/r2_e2e_zero_copy.log
: > $log
for i in 1 2; do
  timeout 200 python profiles/probe_e2e.py >> $log 2>&1
  PP_HOST_NO_ZERO_COPY=1 timeout 200 python profiles/probe_e2e.py >> $log 2>&1
done
PP_HOST_PROBE=2 timeout 200 python profiles/probe_e2e.py >> $log 2>&1
grep -v NCCL $log
timeout 600 python bench.py > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
tail -c 600 gpurun_out/bench_split.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_split.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"]); print(json.dumps(d["e2e"], indent=1))
PY
