#!/bin/bash
# Occupancy sweep (run on the GPU box): rebuild with different min-blocks-per-SM and time the pipeline.
# args per config: PREP CARS DECIDE(PLAN) EMIT
for cfg in "1 1 1 1" "1 6 1 1" "1 8 1 1" "1 1 4 1" "1 1 1 5" "1 1 1 6" "5 1 1 1"; do
  set -- $cfg
  PP_EXTRA_NVCC_FLAGS="-DPP_PREP_MINB=$1 -DPP_CARS_MINB=$2 -DPP_DECIDE_MINB=$3 -DPP_EMIT_MINB=$4" python carnd-path-planning-project_b200/build.py --force > /dev/null 2>&1
  echo "prep=$1 cars=$2 decide=$3 emit=$4: $(python profiles/probe_overhead.py 0 | tail -2 | head -1)"
done
python carnd-path-planning-project_b200/build.py --force > /dev/null 2>&1
