#!/bin/bash
# the repo arm under torchrun on N GPUs, as the driver launches it
N=${1:-8}; tag=${2:-x}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_${tag}_n${N}.json 2> gpurun_out/bench_${tag}_n${N}.err; tail -3 gpurun_out/bench_${tag}_n${N}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}_n${N}.json").read().strip().split("\n")[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["mean_value"], d["e2e"]["whole_rows"]["value"], d["stats"]["check"])
PY
