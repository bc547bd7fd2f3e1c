#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout 300 -p no:cacheprovider tests/test_server_gpu.py > gpurun_out/t_srv.log 2>&1; echo "server tests rc=$?" > gpurun_out/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:decode_attention_mma -s 70 -c 2 -o gpurun_out/prof_r1_decode_attn_mma -f python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-micro --no-graphs > gpurun_out/ncu_attn.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/t_srv.log; tail -2 gpurun_out/ncu_attn.log
