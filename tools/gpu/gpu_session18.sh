#!/bin/bash
# round-2 GPU session 18 (2 GPUs): dry run of the 70B mixed-batch tool on a truncated stack; decode launch list at N=1
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29906 tools/bench_70b_mixed.py --layers 6 --requests 12 > gpurun_out/s18_70b_l6.json 2> gpurun_out/s18_70b_l6.err
echo "rc=$?"; tail -c 1200 gpurun_out/s18_70b_l6.json; tail -5 gpurun_out/s18_70b_l6.err
timeout 300 $TR --master-port 29907 tools/bench_70b_mixed.py --layers 6 --requests 12 --kv-frac 0.35 --rate 0 > gpurun_out/s18_70b_l6_oversub.json 2> gpurun_out/s18_70b_l6_oversub.err
echo "rc=$?"; tail -c 600 gpurun_out/s18_70b_l6_oversub.json; tail -5 gpurun_out/s18_70b_l6_oversub.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 900 --csv --log-file gpurun_out/s18_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-micro --streams 0 --no-ref-shape > gpurun_out/s18_ncu_b.log 2>&1
python tools/launches.py gpurun_out/s18_launches.csv > gpurun_out/s18_launches_summary.txt 2>&1
head -16 gpurun_out/s18_launches_summary.txt
