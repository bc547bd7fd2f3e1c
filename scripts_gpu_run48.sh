#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=$(pwd)
mkdir -p gpurun_out
K="regex:gemm_tc2_kernel|prefill_attention|rope_append|rmsnorm|embed"
( cd $R/_ab_old && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 900 --csv --log-file $R/gpurun_out/launches_old.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> $R/gpurun_out/ncu_old.err )
( cd $R && timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 900 --csv --log-file $R/gpurun_out/launches_new.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> $R/gpurun_out/ncu_new.err )
wc -l gpurun_out/launches_old.csv gpurun_out/launches_new.csv
