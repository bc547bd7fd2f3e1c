#!/bin/bash
# round-2 GPU session 20 (1 GPU): final library — op, engine and server suites
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_server_gpu.py tests/test_weights.py tests/test_gguf.py -q -p no:cacheprovider -m gpu -x > gpurun_out/s20_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/s20_suite.log; tail -3 gpurun_out/s20_suite.log
