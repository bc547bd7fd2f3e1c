#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/batch_decode_profile.py > gpurun_out/batch_profile.txt 2> gpurun_out/batch_profile.err; echo "rc=$?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/batch_profile.txt; tail -n 5 gpurun_out/batch_profile.err
