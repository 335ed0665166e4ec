#!/bin/bash
# Time prebuilt library variants (build/variants/*.so, made on the CPU box with PP_EXTRA_NVCC_FLAGS)
# on the GPU box: per-phase times with the kernels one after the other, then the 4-pipe total.
lib=carnd-path-planning-project_b200/libpp_b200.so
cp $lib /tmp/base.so
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so $lib; else cp build/variants/$v.so $lib; fi
  echo "$v: $(PP_PIPES=1 python profiles/probe_overhead.py 0 1048576 2>&1 | tail -2 | head -1)"
  echo "$v: $(python profiles/probe_overhead.py 0 1048576 2>&1 | tail -1)"
done
cp /tmp/base.so $lib
