#!/bin/bash
mkdir -p gpurun_out
log=gpurun_out/r2_e2e_zc_torch.log
: > $log
PIN=torch timeout 200 python profiles/probe_e2e.py >> $log 2>&1
PIN=torch PP_HOST_NO_ZERO_COPY=1 timeout 200 python profiles/probe_e2e.py >> $log 2>&1
PIN=pp timeout 200 python profiles/probe_e2e.py >> $log 2>&1
grep -v NCCL $log
PP_HOST_NO_ZERO_COPY=1 timeout 600 python bench.py --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('bench no-zero-copy', d['e2e']['value'], d['e2e']['whole_rows']['value'])"
timeout 600 python bench.py --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('bench zero-copy', d['e2e']['value'], d['e2e']['whole_rows']['value'])"
