#!/bin/bash
# Final ncu capture of round 2 (pipeline at 12 cars, final code), after a plain run of the same command.
mkdir -p gpurun_out
K='k_prep|k_cars|k_decide_t|k_fallback|k_emit|k_slow|::stats_kernel'
PP_PIPES=1 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --frames 262144 > /dev/null 2>&1 || exit 1
PP_PIPES=1 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 21 -c 7 -o gpurun_out/prof_r2f python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --frames 262144 > gpurun_out/ncu_r2f.log 2>&1; tail -2 gpurun_out/ncu_r2f.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > /dev/null 2>&1
ls -la gpurun_out/prof_r2f.ncu-rep
