#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/gemm_pair_timeline.py 512 > gpurun_out/pair_timeline.txt 2>&1
timeout 300 python tools/gemm_pair_timeline.py 512 6144 4096 > gpurun_out/pair_timeline_qkv.txt 2>&1
grep "epilogue warp\|span" gpurun_out/pair_timeline.txt gpurun_out/pair_timeline_qkv.txt
