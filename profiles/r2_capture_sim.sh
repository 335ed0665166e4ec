#!/bin/bash
# ncu --set full of the simulator kernels of the rollouts, one stream group so that a launch
# covers all 65,536 rollouts (after a plain run of the same command)
mkdir -p gpurun_out
PP_ROLLOUT_GROUPS=1 python bench.py --workload rollouts --ticks 20 --no-cpu > /dev/null 2>&1 || exit 1
PP_ROLLOUT_GROUPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_sim" -s 40 -c 2 -o gpurun_out/prof_r2j_sim python bench.py --workload rollouts --ticks 20 --no-cpu > gpurun_out/ncu_r2j_sim.log 2>&1; tail -2 gpurun_out/ncu_r2j_sim.log
