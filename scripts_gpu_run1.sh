#!/bin/bash
# first GPU pass: parity tests in increasing risk order, each under its own timeout
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
T="timeout 600 python -m pytest -q -x --timeout 180 -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "not gemm" > gpurun_out/t1_ops.log 2>&1; echo "ops rc=$?" >> gpurun_out/summary.txt
$T tests/test_ops_gpu.py -k "test_gemm and mma" > gpurun_out/t2_gemm_mma.log 2>&1; echo "gemm_mma rc=$?" >> gpurun_out/summary.txt
timeout 600 python -m pytest -q --timeout 180 -p no:cacheprovider tests/test_ops_gpu.py -k "test_gemm and tc" > gpurun_out/t3_gemm_tc.log 2>&1; echo "gemm_tc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest -q --timeout 300 -p no:cacheprovider tests/test_engine_gpu.py -k "mma_sync or chunked or eager" > gpurun_out/t4_engine_mma.log 2>&1; echo "engine_mma rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest -q --timeout 300 -p no:cacheprovider tests/test_engine_gpu.py -k "tcgen05" > gpurun_out/t5_engine_tc.log 2>&1; echo "engine_tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/t6_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t1_ops.log gpurun_out/t2_gemm_mma.log gpurun_out/t3_gemm_tc.log gpurun_out/t4_engine_mma.log gpurun_out/t5_engine_tc.log gpurun_out/t6_smoke.log
cut -c1-1500 gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
