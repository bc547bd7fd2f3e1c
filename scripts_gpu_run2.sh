#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T="timeout 900 python -m pytest -q --timeout 180 -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "not gemm" > gpurun_out/t1_ops.log 2>&1; echo "ops rc=$?" > gpurun_out/summary.txt
$T tests/test_ops_gpu.py -k "test_gemm and mma" > gpurun_out/t2_gemm_mma.log 2>&1; echo "gemm_mma rc=$?" >> gpurun_out/summary.txt
$T tests/test_ops_gpu.py -k "test_gemm and tc" > gpurun_out/t3_gemm_tc.log 2>&1; echo "gemm_tc rc=$?" >> gpurun_out/summary.txt
# launch list of a short bench (1 warm request + 1 timed): per-launch device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-micro --no-graphs > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err; echo "ncu list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/t1_ops.log gpurun_out/t2_gemm_mma.log gpurun_out/t3_gemm_tc.log; do echo "== $f"; tail -n 15 $f; done
wc -l gpurun_out/launches_r1.csv
