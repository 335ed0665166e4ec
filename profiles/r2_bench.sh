#!/bin/bash
# Round-2 measurements on a 1-GPU box: the GPU suite, small-batch latency, the headline bench
# line (and its e2e with whole-row downloads for comparison).
#   gpurun -- 'bash profiles/r2_bench.sh TAG'
tag=${1:-x}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -8 gpurun_out/pytest_$tag.log
timeout 600 python profiles/r2_latency.py 2>&1 | tee gpurun_out/r2_latency_$tag.log
timeout 900 python bench.py > gpurun_out/bench_r2_n1_$tag.json 2> gpurun_out/bench_r2_n1_$tag.err; tail -c 300 gpurun_out/bench_r2_n1_$tag.json; tail -3 gpurun_out/bench_r2_n1_$tag.err
timeout 900 bash profiles/r2_rollouts.sh
timeout 300 python bench.py --workload sweep --no-cpu > gpurun_out/bench_r2_sweep_$tag.json 2>/dev/null; tail -c 300 gpurun_out/bench_r2_sweep_$tag.json
