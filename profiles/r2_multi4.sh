#!/bin/bash
# N-GPU box: PCIe under N-way load, bench weak, dense64 strong.
N=${1:-4}; tag=${2:-x}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 "$@"; }
run profiles/probe_pcie_multi.py 2>/dev/null | tee gpurun_out/pcie_n${N}_$tag.txt
[ "$3" = "pcie-only" ] && exit 0
run bench.py --gpus $N --steps 20 --no-cpu > gpurun_out/bench_r2_n${N}_$tag.json 2> gpurun_out/bench_r2_n${N}_$tag.err; tail -c 300 gpurun_out/bench_r2_n${N}_$tag.json
run bench.py --gpus $N --workload dense64 --no-cpu > gpurun_out/bench_r2_dense64_n${N}_$tag.json 2> gpurun_out/bench_r2_dense64_n${N}_$tag.err; tail -c 300 gpurun_out/bench_r2_dense64_n${N}_$tag.json; tail -2 gpurun_out/bench_r2_dense64_n${N}_$tag.err
