#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2_e2e_sweep.log
: > $log
timeout 200 python profiles/probe_e2e.py >> $log 2>&1
for cap in 65536 131072 262144; do
  for first in 8192 16384 32768 65536; do
    PP_HOST_CHUNK_FIRST=$first PP_HOST_CHUNK_CAP=$cap timeout 200 python profiles/probe_e2e.py >> $log 2>&1
  done
done
grep -v NCCL $log
