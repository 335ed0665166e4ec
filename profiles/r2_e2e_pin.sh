#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2_e2e_pin.log
: > $log
for pin in pp torch pp torch; do
  PIN=$pin timeout 200 python profiles/probe_e2e.py >> $log 2>&1
done
PIN=pp PP_HOST_PROBE=2 timeout 200 python profiles/probe_e2e.py >> $log 2>&1
PIN=pp PP_HOST_CHUNK_FIRST=131072 timeout 200 python profiles/probe_e2e.py >> $log 2>&1
grep -v NCCL $log
