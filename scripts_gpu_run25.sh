#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-micro --no-graphs"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemv_ks_kernel -s 517 -c 4 -o gpurun_out/prof_r1_gemv_ks $B > /dev/null 2> gpurun_out/ncu1.err; echo "ncu gemv rc=$?" > gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none -k "regex:gemv_ks_kernel<1, 3" -s 5 -c 1 -o gpurun_out/prof_r1_lm_head $B > /dev/null 2> gpurun_out/ncu2.err; echo "ncu lm_head rc=$?" >> gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -c 4 -o gpurun_out/prof_r1_gemm_tc2 $B > /dev/null 2> gpurun_out/ncu3.err; echo "ncu tc2 rc=$?" >> gpurun_out/summary.txt
timeout 400 ncu --set full --clock-control none -k regex:decode_attention -s 200 -c 2 -o gpurun_out/prof_r1_decode_attn $B > /dev/null 2> gpurun_out/ncu4.err; echo "ncu attn rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> gpurun_out/ncu5.err; echo "ncu list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1_final.csv
