set -x
python bench.py > gpurun_out/bench_r1m_n1.json 2> gpurun_out/bench_r1m_n1.err
python bench.py --workload rollouts --no-cpu > gpurun_out/bench_r1m_rollouts.json 2>/dev/null
python bench.py --workload sweep --no-cpu > gpurun_out/bench_r1m_sweep.json 2>/dev/null
python bench.py --cars 64 --no-cpu > gpurun_out/bench_r1m_c64.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1m_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
PP_PIPES=1 ncu --set full --clock-control none --import-source on -k regex:"k_|stats_kernel" -s 21 -c 7 -o gpurun_out/prof_r1m python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --frames 262144 > gpurun_out/ncu_r1m.log 2>&1
tail -2 gpurun_out/ncu_r1m.log
python profiles/probe_latency.py 2>&1 | tail -3
