#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python tools/batch_decode_profile.py > gpurun_out/b64_sk.txt 2>&1
LLMLB_GEMM_NO_SK=1 timeout 400 python tools/batch_decode_profile.py > gpurun_out/b64_nosk.txt 2>&1
cat gpurun_out/b64_sk.txt gpurun_out/b64_nosk.txt
