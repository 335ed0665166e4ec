for c in 8 32; do
export CUDA_DEVICE_MAX_CONNECTIONS=$c
for i in 1 2 3; do
python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('conn', os.environ['CUDA_DEVICE_MAX_CONNECTIONS'], 'direct', round(d['value']/1e6,1), round(d['ms_per_step'],1))"
done
PP_ROLLOUT_GRAPH=1 python bench.py --workload rollouts --no-cpu 2>/dev/null | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('conn', os.environ['CUDA_DEVICE_MAX_CONNECTIONS'], 'graph', round(d['value']/1e6,1), round(d['ms_per_step'],1))"
echo "conn $c frames: $(python profiles/probe_overhead.py 0 1048576 2>&1 | tail -1)"
done
