#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 11000 -c 800 --csv --log-file gpurun_out/launches_b64.csv python tools/batch_decode_profile.py > gpurun_out/b64.txt 2> gpurun_out/b64.err; echo "rc=$?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt; wc -l gpurun_out/launches_b64.csv; tail -n 3 gpurun_out/b64.txt
