#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q -x --timeout 300 -p no:cacheprovider tests/test_engine_gpu.py tests/test_server_gpu.py > gpurun_out/t_eng.log 2>&1; echo "engine+server rc=$?" > gpurun_out/summary.txt
timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-micro > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 4 gpurun_out/t_eng.log; cut -c1-120 gpurun_out/bench_loop.json
