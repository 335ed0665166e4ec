#!/bin/bash
# what each part of the host pipeline costs: PP_HOST_PROBE leaves out uploads (1), kernels (2), downloads (4)
mkdir -p gpurun_out
log=gpurun_out/r2_e2e_parts.log
: > $log
for p in 0 2 1 4 3 6 5; do
  echo "PP_HOST_PROBE=$p" >> $log
  PP_HOST_PROBE=$p timeout 200 python profiles/probe_e2e.py >> $log 2>&1
done
grep -v NCCL $log
