#!/bin/bash
# round-2 GPU session 21 (2 GPUs): final library — tp parity (tiny + 8B widths), the 70B tool with Poisson arrivals over the CPU-side group
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 python -m pytest tests/test_tp_gpu.py -q -p no:cacheprovider -k "2-tiny-pairs or 2-8b4-pairs" > gpurun_out/s21_tp.log 2>&1
echo "tp rc=$?" >> gpurun_out/s21_tp.log; tail -2 gpurun_out/s21_tp.log
timeout 120 $TR --master-port 29906 tools/bench_70b_mixed.py --layers 6 --requests 12 --rate 8 > gpurun_out/s21_70b_l6_poisson.json 2> gpurun_out/s21_70b_l6_poisson.err
echo "rc=$?"; tail -c 500 gpurun_out/s21_70b_l6_poisson.json
