#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
R=$(pwd)
mkdir -p gpurun_out
( cd $R/_ab_old && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 330 --csv --log-file $R/gpurun_out/launches_old.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> $R/gpurun_out/ncu_old.err )
( cd $R && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 330 --csv --log-file $R/gpurun_out/launches_new.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-micro --no-graphs > /dev/null 2> $R/gpurun_out/ncu_new.err )
wc -l gpurun_out/launches_old.csv gpurun_out/launches_new.csv
