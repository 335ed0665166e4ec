#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollouts.py -x -q -m gpu > gpurun_out/rollout_tests.log 2>&1; tail -3 gpurun_out/rollout_tests.log
{
for rep in 1 2 3 4 5; do
  echo "auto groups direct: $(timeout 300 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
done
echo "1M rollouts x 60: $(timeout 300 python bench.py --workload rollouts --rollouts 1048576 --ticks 60 --no-cpu 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.1f M ego-frames/s, %.3f ms/tick, launches %d" % (d["value"]/1e6, d["config"]["ms_per_tick"], d["gpu_launches"]))')"
} 2>&1 | tee gpurun_out/r2_rollouts_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_rollouts_chain.csv python bench.py --workload rollouts --ticks 30 --warmup 3 --no-cpu > gpurun_out/ncu_chain.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_rollouts_chain.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
acc = collections.OrderedDict()
for r in rows[hdr + 1:][-160:]:
    nm = r[ki].split("(")[0][-40:]
    acc.setdefault(nm, []).append(float(r[vi].replace(",", "")))
for k, v in acc.items():
    print(f"{k:42s} n={len(v):3d}  mean {sum(v)/len(v)/1e3:8.1f} us")
PY
