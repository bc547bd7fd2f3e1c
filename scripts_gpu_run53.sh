#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests/test_server_gpu.py > gpurun_out/t_srv.log 2>&1; echo "server tests rc=$?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 30 gpurun_out/t_srv.log
