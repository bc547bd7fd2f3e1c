#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 500 --csv --log-file gpurun_out/launches_decode.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-micro --no-graphs > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err; echo "ncu list rc=$?" > gpurun_out/summary.txt
# full capture of the three decode GEMV kernels + attention (3 launches each)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemv_ks_kernel -s 700 -c 6 -o gpurun_out/prof_gemv_ks python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> gpurun_out/ncu_full.err; echo "ncu full gemv rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_attention -s 100 -c 2 -o gpurun_out/prof_decode_attn python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> gpurun_out/ncu_full2.err; echo "ncu full attn rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; ls -la gpurun_out
