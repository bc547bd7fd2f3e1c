#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
LLMLB_PF_MB=0 timeout 300 python tools/decode_timeline.py > gpurun_out/timeline_graph.txt 2> gpurun_out/timeline.err; echo "timeline rc=$?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; cat gpurun_out/timeline_graph.txt; tail -n 5 gpurun_out/timeline.err
