#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2_e2e_same_box.log
nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_pcie_small profiles/probe_pcie_small.cu || exit 1
timeout 300 /tmp/probe_pcie_small > $log 2>&1
PP_HOST_CHUNK_FIRST=131072 timeout 200 python profiles/probe_e2e.py >> $log 2>&1
PP_HOST_CHUNK_FIRST=131072 PP_HOST_PROBE=2 timeout 200 python profiles/probe_e2e.py >> $log 2>&1
timeout 300 /tmp/probe_pcie_small >> $log 2>&1
nproc >> $log; lscpu | grep -i "numa\|model name\|socket" >> $log
nvidia-smi topo -m >> $log 2>&1
grep -v NCCL $log
