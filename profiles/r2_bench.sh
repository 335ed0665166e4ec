#!/bin/bash
# Round-2 measurements on a 1-GPU box: the GPU suite, the headline bench line, configs[4] at its
# single-GPU slice, the C++ multi-device driver on one device.
#   gpurun -- 'bash profiles/r2_bench.sh TAG'
tag=${1:-x}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -6 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_r2_n1_$tag.json 2> gpurun_out/bench_r2_n1_$tag.err; tail -c 600 gpurun_out/bench_r2_n1_$tag.json; tail -3 gpurun_out/bench_r2_n1_$tag.err
python bench.py --workload dense64 --no-cpu > gpurun_out/bench_r2_dense64_n1_$tag.json 2> gpurun_out/bench_r2_dense64_n1_$tag.err; tail -c 400 gpurun_out/bench_r2_dense64_n1_$tag.json; tail -3 gpurun_out/bench_r2_dense64_n1_$tag.err
g++ -std=c++11 -O2 -I include tools/pp_multi.cpp -L carnd-path-planning-project_b200 -lpp_b200 -Wl,-rpath,$PWD/carnd-path-planning-project_b200 -pthread -o /tmp/pp_multi && /tmp/pp_multi data/highway_map.csv --frames 4194304 --steps 5 2>&1 | tee gpurun_out/pp_multi_n1_$tag.json
