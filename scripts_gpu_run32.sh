#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/gemm_decode_bench.py 16 64 > gpurun_out/gemm_decode_bench.txt 2>&1
cat gpurun_out/gemm_decode_bench.txt
