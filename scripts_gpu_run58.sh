#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gemm_tc_kernel|decode_attention|rmsnorm|rope|sample|embed|decode_prepare|gather" -s 3000 -c 900 --csv --log-file gpurun_out/launches_b64_final.csv python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> gpurun_out/ncu_b64.err; echo "rc=$?"
wc -l gpurun_out/launches_b64_final.csv
