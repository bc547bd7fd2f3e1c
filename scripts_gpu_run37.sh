#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/gemm_pair_timeline.py 512 > gpurun_out/pair_timeline.txt 2>&1
cat gpurun_out/pair_timeline.txt
