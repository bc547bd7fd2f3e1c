#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 -k regex:gemm_tc2 -s 2 -c 1 -o gpurun_out/prof_tc2_qkv -f python tools/one_gemm.py 512 6144 4096 0 0 > gpurun_out/ncu_qkv.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_qkv.log; ls -la gpurun_out/prof_tc2_qkv.ncu-rep
