#!/bin/bash
# Round-2 multi-GPU measurements on an N-GPU box:
#   gpurun --gpus N -- 'bash profiles/r2_multi.sh N TAG'
# headline workload (weak scaling, as the driver runs it), configs[4] dense64 (strong scaling, 64M
# frames x 64 cars in total), and the C++ driver with --check at both car counts.
N=${1:-2}; tag=${2:-x}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 20 --no-cpu > gpurun_out/bench_r2_n${N}_$tag.json 2> gpurun_out/bench_r2_n${N}_$tag.err; tail -c 300 gpurun_out/bench_r2_n${N}_$tag.json; tail -2 gpurun_out/bench_r2_n${N}_$tag.err
run --workload dense64 --no-cpu > gpurun_out/bench_r2_dense64_n${N}_$tag.json 2> gpurun_out/bench_r2_dense64_n${N}_$tag.err; tail -c 300 gpurun_out/bench_r2_dense64_n${N}_$tag.json; tail -2 gpurun_out/bench_r2_dense64_n${N}_$tag.err
run --scaling strong --steps 20 --no-cpu > gpurun_out/bench_r2_strong_n${N}_$tag.json 2> gpurun_out/bench_r2_strong_n${N}_$tag.err; tail -c 200 gpurun_out/bench_r2_strong_n${N}_$tag.json
g++ -std=c++11 -O2 -I include tools/pp_multi.cpp -L carnd-path-planning-project_b200 -lpp_b200 -Wl,-rpath,$PWD/carnd-path-planning-project_b200 -pthread -o /tmp/pp_multi
{ timeout 300 /tmp/pp_multi data/highway_map.csv --gpus 1 --frames 8388608 --steps 5
  timeout 300 /tmp/pp_multi data/highway_map.csv --gpus $N --frames 8388608 --steps 5 --check
  timeout 300 /tmp/pp_multi data/highway_map.csv --gpus 1 --frames 8388608 --cars 64 --steps 3
  timeout 300 /tmp/pp_multi data/highway_map.csv --gpus $N --frames 8388608 --cars 64 --steps 3 --check
} 2>&1 | tee gpurun_out/pp_multi_n${N}_$tag.log
