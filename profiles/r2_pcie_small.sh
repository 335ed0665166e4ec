#!/bin/bash
mkdir -p gpurun_out
nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_pcie_small profiles/probe_pcie_small.cu || exit 1
timeout 300 /tmp/probe_pcie_small > gpurun_out/r2_pcie_small.log 2>&1
cat gpurun_out/r2_pcie_small.log
