#!/bin/bash
# ncu --set full of the simulator kernels of the rollouts (after they were spread over eight
# threads per rollout), one stream group so that a launch covers all 65,536 rollouts; plus the
# whole GPU suite once more
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2i.log 2>&1; tail -3 gpurun_out/pytest_r2i.log
PP_ROLLOUT_GROUPS=1 python bench.py --workload rollouts --ticks 20 --no-cpu > /dev/null 2>&1 || exit 1
PP_ROLLOUT_GROUPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sim" -s 40 -c 2 -o gpurun_out/prof_r2i_sim python bench.py --workload rollouts --ticks 20 --no-cpu > gpurun_out/ncu_r2i_sim.log 2>&1; tail -2 gpurun_out/ncu_r2i_sim.log
