#!/bin/bash
# Round-2 measurements on a 1-GPU box: the GPU suite, small-batch latency, the headline bench
# line (and its e2e with whole-row downloads for comparison).
#   gpurun -- 'bash profiles/r2_bench.sh TAG'
tag=${1:-x}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -8 gpurun_out/pytest_$tag.log
timeout 600 python profiles/r2_latency.py 2>&1 | tee gpurun_out/r2_latency_$tag.log
timeout 900 python bench.py > gpurun_out/bench_r2_n1_$tag.json 2> gpurun_out/bench_r2_n1_$tag.err; tail -c 300 gpurun_out/bench_r2_n1_$tag.json; tail -3 gpurun_out/bench_r2_n1_$tag.err
PP_HOST_PITCHED=0 timeout 900 python bench.py --no-cpu --steps 5 > gpurun_out/bench_r2_n1_flatrows_$tag.json 2>/dev/null
python - $tag <<'PY'
import json, sys
for f in ("gpurun_out/bench_r2_n1_TAG.json", "gpurun_out/bench_r2_n1_flatrows_TAG.json"):
    try:
        d = json.loads(open(f.replace("TAG", sys.argv[1])).read().strip().splitlines()[-1])
        print(f, "value %.1f M, e2e %.1f M, d2h %d" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["e2e"]["d2h_bytes_per_step"]))
    except Exception as e:
        print(f, "failed", e)
PY
