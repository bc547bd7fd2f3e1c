#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== narrow prefetch" > gpurun_out/gemm_decode_bench2.txt
timeout 300 python tools/gemm_decode_bench.py 16 64 >> gpurun_out/gemm_decode_bench2.txt 2>&1
echo "== wide prefetch" >> gpurun_out/gemm_decode_bench2.txt
LLMLB_GEMM_PF_WIDE=1 timeout 300 python tools/gemm_decode_bench.py 16 64 >> gpurun_out/gemm_decode_bench2.txt 2>&1
cat gpurun_out/gemm_decode_bench2.txt
