#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout 300 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "gpu suite rc=$?" > gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 6 gpurun_out/t_all.log; tail -2 gpurun_out/smoke.log
