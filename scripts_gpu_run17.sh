#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TRACE_OUT=gpurun_out/trace.npy LLMLB_PF_MB=0 timeout 300 python tools/decode_timeline.py > gpurun_out/timeline.txt 2> gpurun_out/timeline.err; echo "timeline rc=$?" > gpurun_out/summary.txt
ls -la gpurun_out/trace.npy
