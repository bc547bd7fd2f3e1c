#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/tp_plan_check.py > gpurun_out/tp_plan.log 2>&1; echo "tp_plan rc=$?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -v "OMP_NUM_THREADS\|^\*\*\*\*\|^$" gpurun_out/tp_plan.log | tail -n 25
