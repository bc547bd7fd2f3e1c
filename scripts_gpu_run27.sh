#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest -q -rs --timeout 600 -p no:cacheprovider tests/test_engine_gpu.py -m gpu -k "70b" ) > gpurun_out/t_new.log 2>&1; echo "new tests rc=$?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 25 gpurun_out/t_new.log
