#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 4 -o gpurun_out/prof_gemm_tc python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> gpurun_out/ncu_gemm.err; echo "ncu gemm rc=$?" > gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:prefill_attention -c 1 -o gpurun_out/prof_prefill_attn python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2>> gpurun_out/ncu_gemm.err; echo "ncu attn rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; ls -la gpurun_out/*.ncu-rep
