#!/bin/bash
# round-2 GPU session 22 (1 GPU): engine + server suites after the bounded plan-channel back-pressure / failed-engine path
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_engine_gpu.py tests/test_server_gpu.py -q -p no:cacheprovider -m gpu -x --deselect tests/test_engine_gpu.py::test_8b_logits_vs_cpu_oracle > gpurun_out/s22_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/s22_suite.log; tail -3 gpurun_out/s22_suite.log
