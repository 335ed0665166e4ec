#!/bin/bash
# Overlapped step time (4+ chunks in flight) against the share of an SM each kernel's persistent
# grid may take.  Arguments: env settings, one string per configuration.
mkdir -p gpurun_out
{
for cfg in "" "$@"; do
  echo "== [$cfg] $(env $cfg python profiles/probe_overhead.py 2 1048576 2>&1 | tail -1)"
done
} 2>&1 | tee gpurun_out/r2_grids.log
