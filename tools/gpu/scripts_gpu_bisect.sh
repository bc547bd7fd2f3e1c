#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/bisect_8b.py > gpurun_out/bisect.txt 2>&1
cat gpurun_out/bisect.txt | tail -n 20
