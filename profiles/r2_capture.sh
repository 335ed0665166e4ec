#!/bin/bash
# Round-2 ncu captures (1 GPU).  Each capture follows a plain run of the same command that exited 0.
#   gpurun -- 'bash profiles/r2_capture.sh'
mkdir -p gpurun_out
K='k_prep|k_cars|k_decide_t|k_fallback|k_emit|k_slow|::stats_kernel'
set -x
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/pre_launches.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
PP_PIPES=1 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --frames 262144 > /dev/null 2>&1 || exit 1
PP_PIPES=1 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 21 -c 7 -o gpurun_out/prof_r2 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --frames 262144 > gpurun_out/ncu_r2.log 2>&1; tail -2 gpurun_out/ncu_r2.log
PP_PIPES=1 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 21 -c 7 -o gpurun_out/prof_r2_c64 python bench.py --cars 64 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --frames 262144 > gpurun_out/ncu_r2_c64.log 2>&1; tail -2 gpurun_out/ncu_r2_c64.log
PP_PIPES=1 ncu --set full --clock-control none --import-source on -k regex:"k_cars_t|k_decide_t" -s 6 -c 2 -o gpurun_out/prof_r2_v3 python bench.py --variant 3 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --frames 262144 > gpurun_out/ncu_r2_v3.log 2>&1; tail -2 gpurun_out/ncu_r2_v3.log
python bench.py --workload sweep --steps 1 --no-cpu > gpurun_out/bench_r2_sweep.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_sweep" -s 9 -c 3 -o gpurun_out/prof_r2_sweep python bench.py --workload sweep --steps 1 --no-cpu > gpurun_out/ncu_r2_sweep.log 2>&1; tail -2 gpurun_out/ncu_r2_sweep.log
python bench.py --workload rollouts --ticks 20 --no-cpu > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_sim" -s 40 -c 2 -o gpurun_out/prof_r2_sim python bench.py --workload rollouts --ticks 20 --no-cpu > gpurun_out/ncu_r2_sim.log 2>&1; tail -2 gpurun_out/ncu_r2_sim.log
ls -la gpurun_out/*.ncu-rep
