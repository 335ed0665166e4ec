#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/probe_pcie.py > gpurun_out/r2_pcie_pattern.log 2>&1
timeout 300 python profiles/probe_pcie_pattern.py >> gpurun_out/r2_pcie_pattern.log 2>&1
cat gpurun_out/r2_pcie_pattern.log
