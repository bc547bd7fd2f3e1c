#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max > gpurun_out/cpu.txt 2>&1; nproc >> gpurun_out/cpu.txt; python -c "import os;print(len(os.sched_getaffinity(0)))" >> gpurun_out/cpu.txt
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" > gpurun_out/summary.txt
( time timeout 900 python -m pytest -q --timeout 600 -p no:cacheprovider tests/test_engine_gpu.py -k "8b" ) > gpurun_out/t_8b.log 2>&1; echo "8b tests rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/cpu.txt; tail -n 12 gpurun_out/t_8b.log; tail -n 4 gpurun_out/bench_ref.err; cut -c1-700 gpurun_out/bench_ref.json
